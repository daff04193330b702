"""The oracle against the LIVE reference (unmodified /root/reference/models/vit.py) — build container only;
skipped on the GPU box, where the committed fixtures stand in."""
import pytest
import torch

from oracle import reference_shims as rs, vit_oracle, weights as W

pytestmark = pytest.mark.skipif(not rs.reference_available(), reason="/root/reference is not mounted here")


@pytest.fixture(autouse=True)
def _shims():
    yield
    rs.uninstall_shims()


@pytest.mark.parametrize("affine", [True, False])
def test_oracle_equals_reference_vit_tiny(affine):
    sd = W.vit_state_dict("tiny", 64, seed=3, exercise_affine=affine)
    x = W.frames(3, 64, seed=5)
    ref = rs.build_reference_vit("tiny", 64, sd)(x)
    got = vit_oracle.vit_forward(sd, x, 2)
    assert torch.equal(got, ref) or (got - ref).abs().max() < 1e-5


def test_reference_state_dict_schema_is_what_the_dropin_exposes():
    from vidil_b200.blip import create_vit
    ref = rs.build_reference_vit("base", 224)
    ours, width = create_vit("base", 224)
    assert width == 768
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    assert a == b
    assert list(a) == list(b)  # same order too


def test_interpolate_pos_embed_equals_reference():
    from vidil_b200.vision_transformer import VisionTransformer, interpolate_pos_embed
    mod = rs.import_reference_vit()
    enc = VisionTransformer(img_size=384, patch_size=16, embed_dim=128, depth=1, num_heads=2)
    ckpt = torch.randn(1, 14 * 14 + 1, 128, generator=torch.Generator().manual_seed(0))
    assert torch.equal(interpolate_pos_embed(ckpt, enc), mod.interpolate_pos_embed(ckpt, enc))
    assert torch.equal(vit_oracle.interpolate_pos_embed(ckpt, 24 * 24), mod.interpolate_pos_embed(ckpt, enc))


def test_reference_aggregate_function_equals_dropin():
    from vidil_b200.visual_tokenization import aggregate_frame_tokens
    agg = rs.extract_reference_function("run_visual_tokenization.py", 173, 187, "aggregate_frame_tokens")
    frames = [{"objects": ["a", "b", "c"], "attributes": ["x", "y", "x"], "scenes": [], "verbs": ["v", "w", "v"]},
              {"objects": ["b", "a", "d"], "attributes": ["y", "y", "z"], "scenes": [], "verbs": ["w", "w", "u"]}]
    assert aggregate_frame_tokens(frames) == agg(frames)
