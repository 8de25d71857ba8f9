"""-m gpu: candidate-selection indices END TO END (BASELINE.json: "candidate-selection indices are
bit-exact").  Eight candidates (seeds from the reference's get_noises protocol) are denoised by the
CPU oracle (= the reference's arithmetic) and by the CUDA loop through `generate()`; both sets of
final latents go through the same stub verifier, the reference's sort key and top-k rule
(tts_reflectionflow.py:165-182).  The ranking — not only the selected set — must agree exactly.
Entry A (noise search) and entry B (condition stream + LoRA, exact and merged modes)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cases as C  # noqa: E402
from oracle import flux_oracle as fo  # noqa: E402
from reflectionflow_b200.transformer import B200FluxTransformer2DModel  # noqa: E402
from reflectionflow_b200.tts import search as S  # noqa: E402
from reflectionflow_b200.tts.utils import get_noises  # noqa: E402
from reflectionflow_b200.tts.verifiers import Candidate, StubVerifier  # noqa: E402

N_CAND, STEPS = 8, 8


def _rank(latents, seeds, verifier_name):
    cands = [Candidate(f"mid/1_round@{s}.png", s, latents=l) for s, l in zip(seeds, latents)]
    ver = StubVerifier(verifier_name)
    outs = ver.score(cands, ["p"] * len(cands))
    sorted_list = S.sort_outputs(outs, verifier_name, "overall_score")
    topk_idx, _, _ = S.select_topk(outs, sorted_list, cands, len(cands))
    vals = [ver.value(c) for c in cands]
    return topk_idx, outs, vals


# seeds of the global RNG chosen (on the CPU oracle) so that the eight candidates' scores are not
# accidentally within bf16 noise of each other: min gap 1.3e-3 (A, seed 21) / 1.0e-3 (B, seed 9)
@pytest.mark.parametrize("case_name,mode,rng_seed", [("denoiseA_28", None, 21), ("denoiseB_28", "exact", 9),
                                                     ("denoiseB_28", "merged", 9)])
def test_selection_indices_match_the_oracle_loop(case_name, mode, rng_seed):
    case = C.CASES[case_name]
    model, lora = C.build_model(case)
    x = C.build_inputs(case)
    torch.manual_seed(rng_seed)
    noises = get_noises(S.MAX_SEED, N_CAND, case.height, case.width)
    seeds = list(noises)
    m = B200FluxTransformer2DModel(case.config(), lora_rank=case.lora_rank or 0)
    m.load_state_dict(model.state_dict())
    if lora:
        m.load_lora(lora, mode=mode)
    ts, sig = fo.flow_match_sigmas(STEPS, case.n_img)
    ls = fo.LoraSet(lora, 1.0) if lora else None
    ours, ref = [], []
    for s in seeds:
        lat = noises[s]
        o = m.denoise(lat, x["prompt_embeds"], x["pooled"], ts.to(torch.bfloat16) / 1000, sig, case.guidance,
                      x["img_ids"], x["txt_ids"], x["cond_latents"], x["cond_ids"], case.model_config)
        torch.cuda.synchronize()
        ours.append(o.cpu())
        ref.append(fo.denoise(model, lat, x["prompt_embeds"], x["pooled"], STEPS, case.guidance, x["img_ids"],
                              x["txt_ids"], x["cond_latents"], x["cond_ids"], C.oracle_model_config(case), ls))
    m.close()
    for vname in ("nvila", "openai"):
        idx_o, outs_o, vals_o = _rank(ours, seeds, vname)
        idx_r, outs_r, vals_r = _rank(ref, seeds, vname)
        gap = min(abs(a - b) for i, a in enumerate(vals_r) for b in vals_r[i + 1:])
        err = max(abs(a - b) for a, b in zip(vals_o, vals_r))
        print(f"[{case_name} {mode} {vname}] topk_idx {idx_o}; min score gap {gap:.3g}, max |score err| {err:.3g}")
        assert idx_o == idx_r, f"selection differs: ours {idx_o} vs reference arithmetic {idx_r}"
        if vname == "nvila":  # a yes/no label may only differ for a candidate sitting on the p = 0.5 boundary,
            # where "yes, 0.5" and "no, 0.5" take the same place in the reference's sort key
            for o, r, v in zip(outs_o, outs_r, vals_r):
                assert o["label"] == r["label"] or abs(v) < 4 * err
        assert err < 0.25 * gap, "score error is not safely below the smallest gap between candidates"
