"""not-gpu: table-driven pins of the outer loop's selection rules
(tts/tts_reflectionflow.py:152-182,359-448) — candidate-selection indices must be bit-exact."""
import pytest

from reflectionflow_b200.tts import search as S


def nv(name, label, score):
    return {"image_name": name, "label": label, "score": score}


def test_nvila_sort_key_yes_desc_then_no_asc():
    outs = [nv("a", "no", 0.9), nv("b", "yes", 0.6), nv("c", "yes", 0.8), nv("d", "no", 0.55)]
    s = S.sort_outputs(outs, "nvila")
    assert [o["image_name"] for o in s] == ["c", "b", "d", "a"]


def test_openai_sort_desc_and_stability():
    outs = [{"overall_score": {"score": 7}, "i": 0}, {"overall_score": {"score": 9}, "i": 1},
            {"overall_score": {"score": 7}, "i": 2}, {"overall_score": 3, "i": 3}]
    s = S.sort_outputs(outs, "openai", "overall_score")
    assert [o["i"] for o in s] == [1, 0, 2, 3]  # python sort is stable: ties keep input order


def test_topk_indices_and_padding_by_repetition():
    outs = [nv("a", "no", 0.9), nv("b", "yes", 0.6), nv("c", "yes", 0.8)]
    items = ["A", "B", "C"]
    idx, sel, sel_out = S.select_topk(outs, S.sort_outputs(outs, "nvila"), items, 5)
    assert idx[:3] == [2, 1, 0] and sel == ["C", "B", "A", "C", "B"]
    assert [o["image_name"] for o in sel_out] == ["c", "b", "a", "c", "b"]
    idx, sel, _ = S.select_topk(outs, S.sort_outputs(outs, "nvila"), items, 2)
    assert idx == [2, 1] and sel == ["C", "B"]


def test_topk_duplicate_dicts_collapse_like_list_index():
    # equal dicts: outputs.index(x) returns the first one
    same = {"overall_score": {"score": 8, "explanation": "x"}}
    outs = [dict(same), dict(same), {"overall_score": {"score": 2, "explanation": "y"}}]
    idx, sel, _ = S.select_topk(outs, S.sort_outputs(outs, "openai", "overall_score"), ["A", "B", "C"], 3)
    assert idx == [0, 0, 2] and sel == ["A", "A", "C"]


def test_compose_prompts():
    assert S.compose_prompts(["p1", "p2"], ["r1", "r2"]) == ["p1 [Reflexion]: r1", "p2 [Reflexion]: r2"]
    assert S.compose_prompts(["p1"], []) == ["p1"]
    assert S.compose_prompts(["p1"], None) == ["p1"]


def test_chains_round1_and_nvila_break_asymmetry():
    chains = {}
    outs1 = [nv("r1a", "yes", 0.7), nv("r1b", "no", 0.6)]
    S.update_chains(chains, 1, ["r1a", "r1b"], outs1, [], "nvila")
    assert list(chains) == ["r1a", "r1b"] and chains["r1a"]["labels"] == ["yes"]
    # round 2: both children descend from r1a (top-k repeated it)
    outs2 = [nv("r2a", "yes", 0.9), nv("r2b", "no", 0.8)]
    S.update_chains(chains, 2, ["r2a", "r2b"], outs2, ["r1a", "r1a"], "nvila")
    assert chains["r1a"]["images"] == ["r1a", "r2a", "r2b"] and chains["r1b"]["images"] == ["r1b"]
    # a child whose parent sits in TWO chains: nvila stops at the first, openai updates both
    c_nv = {"x": {"images": ["x", "p"], "scores": [1, 2], "labels": ["yes", "yes"]},
            "y": {"images": ["y", "p"], "scores": [1, 2], "labels": ["no", "yes"]}}
    S.update_chains(c_nv, 3, ["child"], [nv("child", "yes", 0.5)], ["p"], "nvila")
    assert c_nv["x"]["images"][-1] == "child" and c_nv["y"]["images"][-1] == "p"
    c_oa = {"x": {"images": ["x", "p"], "scores": [1, 2]}, "y": {"images": ["y", "p"], "scores": [1, 2]}}
    S.update_chains(c_oa, 3, ["child"], [{"m": {"score": 4}}], ["p"], "openai", "m")
    assert c_oa["x"]["images"][-1] == "child" and c_oa["y"]["images"][-1] == "child"


def test_best_per_chain_and_global_best():
    chains = {"a": {"images": ["a0", "a1", "a2"], "labels": ["no", "yes", "yes"], "scores": [0.1, 0.6, 0.9]},
              "b": {"images": ["b0", "b1"], "labels": ["no", "no"], "scores": [0.7, 0.55]}}
    assert S.best_per_chain(chains, "nvila") == ["a2", "b1"]
    assert S.global_best(chains, "nvila") == "a2"
    oa = {"a": {"images": ["a0", "a1"], "scores": [5, 5]}, "b": {"images": ["b0"], "scores": [9]}}
    assert S.best_per_chain(oa, "openai") == ["a0", "b0"]  # np.argmax takes the first maximum
    assert S.global_best(oa, "openai") == "b0"


def test_unknown_verifier_rejected():
    with pytest.raises(NotImplementedError):
        S.sort_outputs([], "gemini")


def test_ours_branch_follows_the_metric_rules():
    """the Image-Verifier ("ours") returns a scalar reward under choice_of_metric"""
    outs = [{"Overall": -0.3}, {"Overall": 1.25}, {"Overall": 0.0}]
    assert [o["Overall"] for o in S.sort_outputs(outs, "ours", "Overall")] == [1.25, 0.0, -0.3]
    ch = S.update_chains({}, 1, ["a", "b", "c"], outs, [], "ours", "Overall")
    ch = S.update_chains(ch, 2, ["a1"], [{"Overall": 0.5}], ["a"], "ours", "Overall")
    assert ch["a"] == {"images": ["a", "a1"], "scores": [-0.3, 0.5]}
    assert S.best_per_chain(ch, "ours") == ["a1", "b", "c"] and S.global_best(ch, "ours") == "b"


def test_records_roundtrip_and_sharding():
    b = S.pack_record(5, 2 ** 31 - 2, 1, 0.123456789)
    assert len(b) == S.RECORD_BYTES and S.unpack_record(b) == (5, 2 ** 31 - 2, 1, 0.123456789)
    assert S.shard_candidates(8, 1, 4) == [1, 5] and S.shard_candidates(3, 2, 4) == [2]
    assert sorted(sum((S.shard_candidates(7, r, 3) for r in range(3)), [])) == list(range(7))


def test_reflection_text_parsers():
    """tts_reflectionflow.py:48-90"""
    from reflectionflow_b200.tts import search as S
    text = ("1. Object:  add a second dog\n- make it brown\n\n"
            "2. Position:  None\n\n"
            "3. Style:  sharper focus\n\nfree text without a title")
    assert S.extract_reflections([text, ""]) == [
        {"Object": ["add a second dog", "make it brown"], "Position": ["None"], "Style": ["sharper focus"]}, {}]
    assert S.concat_extract_reflections([text]) == ["add a second dog make it brownsharper focus"]
    import pytest
    with pytest.raises(IndexError):
        S.concat_extract_reflections(["Title: no number"])
