"""Host-side mirror of the reference interface: factories, state_dict schema, ontology clean-up, token aggregation,
the no-CPU-fallback rule."""
import json
import os

import pytest
import torch

from oracle import tokenization_oracle, weights as W
from vidil_b200 import visual_tokenization as vt
from vidil_b200.blip import create_vit
from vidil_b200.clip import CLIPVisionB200
from vidil_b200.vision_transformer import VisionTransformer


def test_create_vit_large_schema():
    enc, width = create_vit("large", 224)
    assert width == 1024
    sd = enc.state_dict()
    assert len(sd) == 294                                     # SURVEY.md §8b (probed on the reference)
    assert tuple(sd["pos_embed"].shape) == (1, 197, 1024)
    assert tuple(sd["patch_embed.proj.weight"].shape) == (1024, 3, 16, 16)
    assert tuple(sd["blocks.23.attn.qkv.weight"].shape) == (3072, 1024)
    assert tuple(sd["blocks.0.mlp.fc2.weight"].shape) == (1024, 4096)
    assert enc.patch_embed.num_patches == 196
    ref_sd = W.vit_state_dict("large", 224)
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == {k: tuple(v.shape) for k, v in sd.items()}


def test_create_vit_errors_like_the_reference():
    with pytest.raises(ValueError):
        create_vit("huge", 224)
    enc, width = create_vit("base", 384)
    assert width == 768 and enc.pos_embed.shape[1] == 577


def test_load_state_dict_strict_roundtrip():
    sd = W.vit_state_dict("tiny", 32)
    m = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=2, num_heads=2)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["blocks.1.mlp.fc1.bias"], sd["blocks.1.mlp.fc1.bias"])


def test_no_cpu_path():
    m = VisionTransformer(img_size=32, patch_size=16, embed_dim=128, depth=1, num_heads=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 32, 32))
    c = CLIPVisionB200(**W.CLIP_CONFIGS["tiny"])
    with pytest.raises(RuntimeError, match="CUDA"):
        c(torch.zeros(1, 3, 28, 28))
    from vidil_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.sim_topk(torch.zeros(2, 64), torch.zeros(8, 64), 2)


def test_product_never_imports_the_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vidil_b200")
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_clip_state_dict_uses_transformers_names():
    c = CLIPVisionB200(**W.CLIP_CONFIGS["tiny"])
    sd = W.clip_vision_state_dict("tiny")
    assert set(c.state_dict()) == set(sd)
    c.load_state_dict(sd)
    assert torch.equal(c.state_dict()["visual_projection.weight"], sd["visual_projection.weight"])
    names = dict(c._packed_tensors())
    assert tuple(names["blocks.0.attn.qkv.weight"].shape) == (3 * 128, 128)
    assert torch.equal(names["blocks.1.attn.qkv.bias"][128:256], sd["vision_model.encoder.layers.1.self_attn.k_proj.bias"])


def test_clip_text_state_dict_uses_transformers_names():
    from vidil_b200.clip import CLIPTextB200
    c = W.CLIP_TEXT_CONFIGS["tiny"]
    m = CLIPTextB200(**c)
    sd = W.clip_text_state_dict("tiny")
    assert set(m.state_dict()) == set(sd)
    m.load_state_dict(sd)
    names = dict(m._packed_tensors())
    assert tuple(names["token_embedding"].shape) == (c["vocab_size"], 128)
    assert tuple(names["blocks.1.attn.qkv.weight"].shape) == (384, 128)
    ids = torch.tensor([[94, 5, 7, 95, 95], [94, 9, 95, 95, 95]])
    assert m.eos_positions(ids).tolist() == [3, 2]
    with pytest.raises(RuntimeError, match="CUDA"):
        m(ids)


def test_aggregate_matches_fixture(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "tokenization.json")))
    for case in g["aggregate_cases"]:
        assert vt.aggregate_frame_tokens(case["frame_tokens"]) == case["aggregated"]
        assert vt.aggregate_frame_tokens(case["frame_tokens"]) == tokenization_oracle.aggregate_frame_tokens(case["frame_tokens"])


def test_prompt_functions():
    assert vt.get_prefix_prompt_functions("v1")["verbs"]("running") == "A photo of running"
    assert vt.get_prefix_prompt_functions("v0")["objects"]("cat") == "cat"
    assert set(vt.get_prefix_prompt_functions("v1")) == {"objects", "attributes", "scenes", "verbs"}


def test_ontology_cleanup_keeps_the_reference_quirk(tmp_path):
    files = vt.ONTOLOGY_FILES["vg"]
    data = {"objects": ["cat", "dog", "video", "tree"], "attributes": ["cat", "dog", "red", "tree", "blue", "stock"],
            "scenes": ["beach", "audio"], "verbs": {"run": 1, "sound": 2, "jump": 3}}
    for key, rel in files.items():
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(json.dumps(data[key]))
    got = vt.load_ontology("vg", root=str(tmp_path))
    # removing while iterating skips the element after each removal: "dog" survives although it is an object
    assert got["attributes"] == ["dog", "red", "blue"]
    assert got["objects"] == ["cat", "dog", "tree"]
    assert got["scenes"] == ["beach"]
    assert got["verbs"] == ["run", "jump"]
    with pytest.raises(KeyError):
        vt.load_ontology("nope", root=str(tmp_path))


@pytest.mark.skipif(not os.path.isdir("/root/reference/visual_token_ontology"), reason="reference tree not mounted")
def test_real_vg_ontology_sizes():
    got = vt.load_ontology("vg", root="/root/reference")
    assert len(got["scenes"]) == 365
    assert 19000 < len(got["objects"]) < 20000 and 7000 < len(got["verbs"]) < 7500


def test_clip_model_rejects_unsupported_text_tower_instead_of_falling_back():
    """North-star: no library / CPU fallback — a text tower outside the native kernels' shapes is an error."""
    import pytest
    from transformers import CLIPConfig, CLIPModel

    from vidil_b200.clip import VidilCLIPModel
    cfg = CLIPConfig(text_config=dict(hidden_size=96, intermediate_size=192, num_hidden_layers=1, num_attention_heads=2,
                                      max_position_embeddings=16, vocab_size=64),
                     vision_config=dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2,
                                        image_size=28, patch_size=14), projection_dim=64)
    hf = CLIPModel(cfg).eval()
    with pytest.raises(RuntimeError, match="unsupported CLIP text tower"):
        VidilCLIPModel(hf)
    m = VidilCLIPModel(hf, native_text=False)       # image tower only is allowed; asking it for text is an error
    with pytest.raises(RuntimeError, match="no fallback text path"):
        m(input_ids=__import__("torch").zeros(1, 4, dtype=__import__("torch").long))
