"""Developer tool: how closely does the operand-emulating oracle (oracle/med_oracle.emulate) track the native text stack?
Hidden states of the tiny decoder, teacher-forced, with sub-layers switched off one at a time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import med_oracle, weights as W  # noqa: E402
from vidil_b200.med import BertConfig, BertLMHeadModel  # noqa: E402

dev = torch.device("cuda")
name, batch, T, n_img = "tiny", 3, 9, 5
c, sp = W.MED_CONFIGS[name], W.MED_SPECIAL[name]
enc = W.image_tokens(batch, n_img, c["encoder_width"], seed=0)
ids, _ = W.caption_ids(name, batch, T, seed=0, min_words=T - 2)
ids[:, 0] = sp["bos"]


def variant(sd, off):
    sd = {k: v.clone() for k, v in sd.items()}
    for k in sd:
        if any(o in k for o in off) and ("dense.weight" in k or "dense.bias" in k):
            sd[k].zero_()
    return sd


base = W.med_state_dict(name, "decoder", seed=0)
cases = {"all on": [], "self only": ["crossattention.output", ".output.dense", "intermediate"],
         "cross only": ["attention.output.dense", ".output.dense", "intermediate"], "ffn only": ["attention.output", "crossattention.output"]}
for dtype, tdt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
    for label, off in cases.items():
        if label == "self only":
            sd = variant(base, ["crossattention.output.dense", "intermediate.dense"])
            for k in sd:
                if ".output.dense" in k and "attention" not in k:
                    sd[k].zero_()
        elif label == "cross only":
            sd = variant(base, ["intermediate.dense"])
            for k in sd:
                if (".attention.output.dense" in k) or (".output.dense" in k and "attention" not in k):
                    sd[k].zero_()
        elif label == "ffn only":
            sd = variant(base, ["attention.output.dense", "crossattention.output.dense"])
        else:
            sd = base
        m = BertLMHeadModel(BertConfig(**c), compute_dtype=dtype)
        m.load_state_dict({k[len("text_decoder."):]: v for k, v in sd.items()}, strict=False)
        m = m.to(dev).eval()
        hid, _, _ = m.bert.run(ids, None, enc.to(dev), causal=True, want_hidden=True)
        hid = hid.cpu()
        with torch.no_grad():
            r32, _ = med_oracle.bert_forward(sd, "text_decoder.bert.", ids, None, enc, c["num_attention_heads"], c["num_hidden_layers"], causal=True)
            with med_oracle.emulate(tdt):
                re_, _ = med_oracle.bert_forward(sd, "text_decoder.bert.", ids, None, enc, c["num_attention_heads"], c["num_hidden_layers"], causal=True)
        print(f"{dtype} {label:10s}: native-fp32 {float((hid - r32).abs().max()):.3e}  native-emulated {float((hid - re_).abs().max()):.3e}  "
              f"emulated-fp32 {float((re_ - r32).abs().max()):.3e}", flush=True)
