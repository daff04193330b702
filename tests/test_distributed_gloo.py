"""world_size-2 run of the N>1 path on CPU (gloo): reference partition, one all-gather of JSON rows, rank-ordered
merge, rank-0 file identical to what the reference's tmp-file merge would write."""
import json
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, n_videos):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist

    from vidil_b200 import distributed as vd
    info = vd.init_distributed_mode(backend="gloo")
    assert info["distributed"] and vd.get_world_size() == world and vd.get_rank() == rank
    videos = [f"video{i}" for i in range(n_videos)]
    s, e = vd.shard_bounds(len(videos))
    rows = {v: {"frame_tokens": [{"objects": [f"obj{rank}", "é∑"]}], "caption": [f"caption of {v}"], "rank": rank}
            for v in videos[s:e]}
    merged = vd.gather_and_write(rows, os.path.join(out_dir, "visual_tokens.json"))
    if rank == 0:
        assert list(merged) == videos          # rank order == original order for contiguous slices
    else:
        assert merged is None
    # every rank sees every rank's object from all_gather_json
    got = vd.all_gather_json({"r": rank, "empty": {}})
    assert [g["r"] for g in got] == list(range(world))
    # the phrase bank embedded once ACROSS the ranks: every rank embeds a contiguous run of the 512-phrase batches, the rows
    # come back in the original order on every rank (all_gather_rows), including ranks that got no batch at all
    import torch
    from types import SimpleNamespace

    from vidil_b200 import visual_tokenization as vt
    texts = [f"phrase {i}" for i in range(1300)]                      # 3 batches of <= 512

    class Proc:
        def __call__(self, text=None, **kw):
            ids = torch.tensor([[len(t), int(t.split()[1])] for t in text])
            return {"input_ids": ids, "attention_mask": torch.ones_like(ids)}

    class Model:
        def __call__(self, input_ids=None, attention_mask=None):
            return SimpleNamespace(text_embeds=torch.stack([input_ids[:, 1].float(), input_ids[:, 1].float() * 0.5], dim=1))

    whole, _, _ = vt.get_text_embeddings_clip(Model(), Proc(), texts, "cpu")
    sharded, _, _ = vt.get_text_embeddings_clip(Model(), Proc(), texts, "cpu", shard_over_ranks=True)
    assert torch.equal(whole, sharded) and tuple(sharded.shape) == (1300, 2)
    assert torch.equal(vd.all_gather_rows(torch.full((rank, 3), float(rank))),
                       torch.cat([torch.full((r, 3), float(r)) for r in range(world)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_videos", [(2, 7), (2, 1), (3, 10)])
def test_gloo_shard_gather_merge(tmp_path, world, n_videos):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), n_videos), nprocs=world, join=True)
    merged = json.load(open(tmp_path / "visual_tokens.json"))
    assert list(merged) == [f"video{i}" for i in range(n_videos)]
    # what the reference's file merge would have produced: dict.update in rank order, json.dump(indent=4)
    expect = {}
    for rank in range(world):
        step = n_videos // world + 1
        for i in range(rank * step, min(n_videos, rank * step + step)):
            expect[f"video{i}"] = {"frame_tokens": [{"objects": [f"obj{rank}", "é∑"]}], "caption": [f"caption of video{i}"],
                                   "rank": rank}
    assert merged == expect
    assert open(tmp_path / "visual_tokens.json").read() == json.dumps(expect, indent=4)


def test_single_process_fallbacks():
    from vidil_b200 import distributed as vd
    assert vd.get_world_size() == 1 and vd.get_rank() == 0 and vd.is_main_process()
    assert vd.shard_bounds(10) == (0, 10)
    assert vd.all_gather_json({"a": 1}) == [{"a": 1}]
    assert vd.merge_rank_dicts([{"a": 1}, None, {"a": 2, "b": 3}]) == {"a": 2, "b": 3}
    import torch
    t = torch.arange(6.).view(3, 2)
    assert vd.all_gather_rows(t) is t


def test_shard_bounds_matches_fixture(golden_dir):
    from vidil_b200 import distributed as vd
    for case in json.load(open(os.path.join(golden_dir, "sharding.json"))):
        for rank, sl in enumerate(case["slices"]):
            s, e = vd.shard_bounds(case["n"], case["world"], rank)
            assert ([s, e] if e > s else None) == sl
