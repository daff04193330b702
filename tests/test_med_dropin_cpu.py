"""Drop-in surface of the CapFilt models without a GPU: the parameter trees of vidil_b200.med / vidil_b200.blip carry exactly the
reference's state_dict keys and shapes (so BLIP checkpoints load with no missing key, models/blip.py:273,352), and
load_checkpoint keeps the reference's behaviour (position-embedding interpolation, shape-mismatch dropping)."""
import os

import pytest
import torch

from oracle import reference_shims as rs, weights as W
from vidil_b200.blip import BLIP_Decoder, BLIP_ITM, blip_decoder, blip_itm, load_checkpoint
from vidil_b200.med import BertConfig, BertLMHeadModel, BertModel


def _ref_keys(model):
    return {k: tuple(v.shape) for k, v in model.state_dict().items()}


@pytest.mark.skipif(not rs.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("name", ["tiny", "base_b"])
def test_med_parameter_tree_equals_the_reference_classes(name):
    c = W.MED_CONFIGS[name]
    sd = W.med_state_dict(name, "decoder", seed=0)
    ref_dec, _ = rs.build_reference_med(name, "decoder", sd)
    ours = BertLMHeadModel(BertConfig(**c))
    ref, mine = _ref_keys(ref_dec), _ref_keys(ours)
    # the reference registers cls.predictions.decoder.bias as an alias of cls.predictions.bias (med.py:538); nothing else differs
    assert set(ref) - set(mine) <= {"cls.predictions.decoder.bias"}
    assert set(mine) - set(ref) <= {"bert.embeddings.position_ids"}        # a non-persistent buffer in newer transformers
    assert all(mine[k] == ref[k] for k in mine if k in ref)
    # a reference state_dict loads into the drop-in with strict=True (the alias is dropped by the load hook) and the values arrive
    ref_sd = ref_dec.state_dict()
    missing, unexpected = ours.load_state_dict(ref_sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing)
    assert torch.equal(ours.cls.predictions.bias, ref_sd["cls.predictions.bias"])
    assert torch.equal(ours.bert.encoder.layer[1].crossattention.self.key.weight,
                       ref_sd["bert.encoder.layer.1.crossattention.self.key.weight"])
    # a checkpoint of a weight-tied model that stores the word embeddings only still fills the output matrix
    tied = {k: v for k, v in ref_sd.items() if k != "cls.predictions.decoder.weight"}
    ours2 = BertLMHeadModel(BertConfig(**c))
    ours2.load_state_dict(tied, strict=False)
    assert torch.equal(ours2.cls.predictions.decoder.weight, ref_sd["bert.embeddings.word_embeddings.weight"])

    sdi = W.med_state_dict(name, "itm", seed=0)
    ref_enc, _ = rs.build_reference_med(name, "itm", sdi)
    enc = BertModel(BertConfig(**c))
    ref, mine = _ref_keys(ref_enc), _ref_keys(enc)
    assert set(ref) ^ set(mine) <= {"embeddings.position_ids"}
    assert all(mine[k] == ref[k] for k in mine if k in ref)


def test_blip_models_load_a_blip_style_checkpoint(tmp_path):
    """blip_decoder / blip_itm(pretrained=...) as run_video_CapFilt.py:143,148 call them: a checkpoint trained at 384 px loads
    into a 224 px model through the reference's position-embedding interpolation, with no missing key."""
    src = BLIP_Decoder(image_size=384, vit="base", prompt_ids=[101, 1037, 3861, 1997, 102])
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    sd["text_decoder.cls.predictions.decoder.bias"] = sd["text_decoder.cls.predictions.bias"]      # as BLIP checkpoints carry it
    path = os.path.join(tmp_path, "model_base_capfilt.pth")
    torch.save({"model": sd}, path)
    m = blip_decoder(pretrained=path, image_size=224, vit="base", prompt_ids=[101, 1037, 3861, 1997, 102])
    assert m.visual_encoder.pos_embed.shape == (1, 197, 768)
    assert torch.equal(m.text_decoder.bert.encoder.layer[3].intermediate.dense.weight,
                       src.text_decoder.bert.encoder.layer[3].intermediate.dense.weight)
    assert torch.equal(m.visual_encoder.pos_embed[:, 0], src.visual_encoder.pos_embed[:, 0])        # class-token row kept
    assert m.prompt_length == 4 and m.bos_token_id == 30522 and m.sep_token_id == 102

    isrc = BLIP_ITM(image_size=224, vit="base")
    ipath = os.path.join(tmp_path, "model_base_retrieval.pth")
    isd = {k: v.clone() for k, v in isrc.state_dict().items()}
    isd["temp"] = torch.tensor(0.07)                                       # retrieval checkpoints carry extra tensors: ignored
    torch.save({"model": isd}, ipath)
    f = blip_itm(pretrained=ipath, image_size=224, vit="base")
    assert torch.equal(f.itm_head.weight, isrc.itm_head.weight)
    assert f.text_encoder._cls_head[0] is f.itm_head
    with pytest.raises(RuntimeError):
        load_checkpoint(f, os.path.join(tmp_path, "missing.pth"))
    with pytest.raises(RuntimeError):                                      # CPU module: there is no CPU path
        f(torch.zeros(1, 3, 224, 224), ["a caption"])
