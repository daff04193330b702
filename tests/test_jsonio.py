"""vidil_b200.jsonio writes the reference's result-file format (json.dump(..., indent=4)) byte for byte."""
import json
import random
import time

import pytest

from vidil_b200 import jsonio


def _random_obj(rng, depth=0):
    kinds = ["str", "int", "float", "bool", "none", "list", "dict", "strlist", "tuple"]
    k = rng.choice(kinds if depth < 4 else kinds[:5])
    if k == "str":
        return "".join(rng.choice(['a', 'Z', ' ', '"', '\\', '\n', '\t', '\x07', 'é', '漢', '😀', '/', '\x7f', "'"]) for _ in range(rng.randint(0, 12)))
    if k == "int":
        return rng.choice([0, -1, 7, 10 ** 20, -(2 ** 63)])
    if k == "float":
        return rng.choice([0.0, -0.0, 1.5, 1e-7, 1e22, 0.1 + 0.2, float("inf"), float("-inf"), float("nan"), 3.0])
    if k == "bool":
        return rng.random() < 0.5
    if k == "none":
        return None
    if k == "strlist":
        return [f"phrase {i} \"q\"" for i in range(rng.randint(0, 5))]
    if k == "list":
        return [_random_obj(rng, depth + 1) for _ in range(rng.randint(0, 4))]
    if k == "tuple":
        return tuple(_random_obj(rng, depth + 1) for _ in range(rng.randint(0, 3)))
    keys = [rng.choice(["video1", "k\"ey", "é", 3, 2.5, True, False, None, "", float("inf")]) for _ in range(rng.randint(0, 4))]
    return {key: _random_obj(rng, depth + 1) for key in keys}


def test_dumps_indent4_is_byte_identical_to_the_standard_library():
    rng = random.Random(0)
    for _ in range(3000):
        obj = _random_obj(rng)
        assert jsonio.dumps_indent4(obj) == json.dumps(obj, indent=4)


def test_result_file_shapes_and_subclasses():
    tokens = {f"video{i}": {"objects": [f"a thing {j}" for j in range(5)], "scenes": [], "frame_scores": [[0.25, 1], [0.5, 2]],
                            "n": i, "ok": True, "none": None} for i in range(20)}
    assert jsonio.dumps_indent4(tokens) == json.dumps(tokens, indent=4)

    class MyDict(dict):
        pass

    class MyStr(str):
        pass

    nested = {"a": MyDict(x=[1, MyStr("s")]), "b": [MyDict()]}
    assert jsonio.dumps_indent4(nested) == json.dumps(nested, indent=4)
    with pytest.raises(TypeError):
        jsonio.dumps_indent4({"a": object()})
    with pytest.raises(TypeError):
        jsonio.dumps_indent4({(1, 2): 3})


def test_dump_matches_on_a_result_file_and_is_not_slower(tmp_path):
    tokens = {f"video{i}": {t: [f"a photo of thing number {j}" for j in range(15)] for t in ("objects", "attributes", "scenes", "verbs")}
              for i in range(3000)}
    ref = json.dumps(tokens, indent=4)
    with open(tmp_path / "x.json", "w") as f:
        jsonio.dump_indent4(tokens, f)
    assert open(tmp_path / "x.json").read() == ref

    def best_of(fn, n=3):
        best = float("inf")
        for _ in range(n):
            t = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t)
        return best

    t_json, t_fast = best_of(lambda: json.dumps(tokens, indent=4)), best_of(lambda: jsonio.dumps_indent4(tokens))
    print(f"{len(ref) / 1e6:.1f} MB: json {t_json:.3f} s, jsonio {t_fast:.3f} s")
    assert t_fast < 1.5 * t_json      # typically 0.55x; the bound only catches a regression to something pathological
