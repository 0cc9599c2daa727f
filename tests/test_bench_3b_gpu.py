"""PaliGemma-3B architecture (random init) at FULL widths: logits parity with the CPU fp32 oracle on what bench.py times.

  * 3B-224, R1 (well conditioned): B = 1 teacher-forced logits, greedy 32 tokens identical, batch rows independent;
  * 3B-224, R2 (the benchmarked regime), B = 64, S = 260: teacher-forced logits of 4 sampled rows of the 64-row job
    (rows are independent, so B = 1 oracle runs of those rows are the reference), and the top-p draws of the sampled job
    against the oracle's kept set;
  * 3B-448 (1024 image tokens) and 3B-896 (4096 image tokens), B = 1: prefill + 4 teacher-forced decode steps.

Tolerance (north_star: bf16 tolerance stated in the test): last-position logits cosine >= 0.9999 and max-abs <= 1 % of the
oracle logit absmax in R1; cosine >= 0.999 and <= 3 % in the diffuse regime R2; greedy tokens identical to the oracle's
whenever its top-2 margin exceeds 4x the measured error (all 32 steps in R1).  Every measured figure is also written to
gpurun_out/parity_r02.json (copied to profiles/parity_r02.json, which bench.py attaches to its line as `parity`)."""
import json
import os

import pytest
import torch

from parity_utils import build_model, stats, top2_margin

pytestmark = pytest.mark.gpu

from oracle import paligemma_oracle as O  # noqa: E402
from paligemma_multimodal_system_b200.random_init import make_inputs, make_state_dict, paligemma_3b_config  # noqa: E402


OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_r02.json")


def _record(name, payload):
    """Accumulates the measured parity figures of this run (tracked copy: profiles/parity_r02.json)."""
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    try:
        data = json.load(open(OUT))
    except (OSError, ValueError):
        data = {}
    data[name] = payload
    data["tolerance"] = {"R1": "cos >= 0.9999, max-abs <= 1% of oracle absmax", "R2": "cos >= 0.999, max-abs <= 3% of oracle absmax",
                         "greedy": "identical wherever the oracle top-2 margin > 4x max-abs error"}
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _summ(rows):
    """rows: list of stats dicts (+ margin) -> worst-case summary."""
    return {"max_abs_pct_of_absmax": round(100 * max(r["rel"] for r in rows), 4), "max_abs": max(r["max_abs"] for r in rows),
            "min_cosine": round(min(r["cos"] for r in rows), 7), "oracle_absmax": max(r["absmax"] for r in rows),
            "min_oracle_top2_margin": min(r["margin"] for r in rows), "checks": len(rows)}


@pytest.fixture(scope="module")
def setup_3b():
    cfg = paligemma_3b_config(224)
    sd = make_state_dict(cfg, "R1", seed=0, device="cuda", dtype=torch.bfloat16)
    model = build_model(cfg, sd)
    sd_cpu = {k: v.float().cpu() for k, v in sd.items()}
    return cfg, model, sd_cpu


def test_3b_prefill_and_teacher_forced_decode_logits(setup_3b):
    cfg, model, sd_cpu = setup_3b
    steps = 6
    inp = make_inputs(cfg, batch=1, prompt_len=4, seed=0)
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    ref_t, ref_l = O.generate(sd_cpu, cfg, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], steps, return_logits=True)
    toks, logits = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), steps,
                                  return_logits=True, forced_tokens=ref_t)
    rows = []
    for t in range(steps):
        s = stats(logits[0, t], ref_l[0, t])
        rows.append(dict(s, margin=float(top2_margin(ref_l[0, t]))))
        print(f"[parity-3B] step {t}: max_abs={s['max_abs']:.4g} ({100 * s['rel']:.3f}% of absmax {s['absmax']:.4g}) cos={s['cos']:.6f} "
              f"oracle margin {float(top2_margin(ref_l[0, t])):.4g}")
        assert s["cos"] >= 0.9999 and s["rel"] <= 0.01, (t, s)
        if float(top2_margin(ref_l[0, t])) > 4 * s["max_abs"]:
            assert int(toks[0, t]) == int(ref_t[0, t])
    _record("3b_224_R1_b1_teacher_forced", dict(_summ(rows), workload="3B-224 R1, B=1, prefill + 5 teacher-forced decode steps"))


def test_3b_greedy_32_tokens_identical_and_batch_rows_independent(setup_3b):
    cfg, model, sd_cpu = setup_3b
    inp = make_inputs(cfg, batch=3, prompt_len=4, seed=0)
    # oracle for row 0 only (CPU cost), 32 free-running greedy tokens
    ref_t, ref_l = O.generate(sd_cpu, cfg, inp["input_ids"][:1], inp["pixel_values"][:1], inp["attention_mask"][:1], 32, return_logits=True)
    toks = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 32)
    m = top2_margin(ref_l[0])
    print(f"[parity-3B] greedy row0 {toks[0].tolist()} oracle {ref_t[0].tolist()} oracle top-2 margin min {m.min():.4g} median {m.median():.4g}")
    assert toks[0].tolist() == ref_t[0].tolist()
    _record("3b_224_R1_greedy32", {"identical_tokens": 32, "of": 32, "oracle_top2_margin_min": float(m.min()),
                                   "oracle_top2_margin_median": float(m.median()), "workload": "3B-224 R1, B=3 job, row 0 vs B=1 oracle run"})
    single = model.generate(inp["input_ids"][:1].cuda(), inp["pixel_values"][:1].cuda(), inp["attention_mask"][:1].cuda(), 32)
    assert single[0].tolist() == toks[0].tolist()


# ---------------------------------------------------------------------------------------------------------------------
# the benchmarked workload (BASELINE configs[2]): R2, 64 requests, S = 260
# ---------------------------------------------------------------------------------------------------------------------
ROWS = (0, 21, 42, 63)


@pytest.fixture(scope="module")
def setup_3b_r2():
    cfg = paligemma_3b_config(224)
    sd = make_state_dict(cfg, "R2", seed=0, device="cuda", dtype=torch.bfloat16)  # bench.py's weights
    model = build_model(cfg, sd)
    sd_cpu = {k: v.float().cpu() for k, v in sd.items()}
    del sd
    inp = make_inputs(cfg, batch=64, prompt_len=4, seed=100)  # bench.py's rank-0 batch
    return cfg, model, sd_cpu, inp


def test_3b_r2_batch64_teacher_forced_rows_vs_oracle(setup_3b_r2):
    """The 64-row job bench.py times, teacher-forced with one fixed token matrix: prefill + 4 decode steps; rows 0 / 21 / 42 / 63
    against B = 1 oracle runs of the same rows (the decode batch must not couple rows: one GEMM N tile, one attention grid)."""
    cfg, model, sd_cpu, inp = setup_3b_r2
    steps = 5
    g = torch.Generator().manual_seed(17)
    forced = torch.randint(3, 250000, (64, steps), generator=g)
    toks, logits = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), steps,
                                  return_logits=True, forced_tokens=forced)
    rows = []
    for r in ROWS:
        ref_t, ref_l = O.generate(sd_cpu, cfg, inp["input_ids"][r:r + 1], inp["pixel_values"][r:r + 1], inp["attention_mask"][r:r + 1],
                                  steps, return_logits=True, forced_tokens=forced[r:r + 1])
        for t in range(steps):
            s = stats(logits[r, t], ref_l[0, t])
            mg = float(top2_margin(ref_l[0, t]))
            rows.append(dict(s, margin=mg))
            print(f"[parity-3B-R2-b64] row {r} step {t}: max_abs={s['max_abs']:.4g} ({100 * s['rel']:.3f}% of absmax {s['absmax']:.4g}) "
                  f"cos={s['cos']:.6f} oracle margin {mg:.4g}")
            assert s["cos"] >= 0.999 and s["rel"] <= 0.03, (r, t, s)
            if mg > 4 * s["max_abs"]:
                assert int(toks[r, t]) == int(ref_t[0, t])
    _record("3b_224_R2_b64_teacher_forced", dict(_summ(rows), rows=list(ROWS),
                                                  workload="3B-224 R2 (bench.py weights / inputs), B=64, S=260, prefill + 4 teacher-forced decode steps"))


def test_3b_r2_batch64_top_p_draws_in_oracle_kept_set(setup_3b_r2):
    """bench.py's sampling job (temperature 0.8, top-p 0.9, 64 rows): every token drawn for rows 0 / 21 / 42 / 63 over 8 steps must
    lie in the top-p kept set of the ORACLE's distribution for that row's history.  A drawn token can sit in the band at the
    kept-set boundary that bf16 logit noise moves (the kept set of the diffuse regime holds thousands of tokens), so
    membership is asserted against the oracle set at top_p + 0.03 and the strict fraction is reported."""
    cfg, model, sd_cpu, inp = setup_3b_r2
    steps, temp, top_p = 8, 0.8, 0.9
    toks = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), steps,
                          do_sample=True, temperature=temp, top_p=top_p, seed=1234).cpu()
    assert toks.shape == (64, steps) and int(toks.min()) >= 0 and int(toks.max()) < cfg["text_config"]["vocab_size"]
    strict = loose = total = 0
    kept_sizes = []
    for r in ROWS:
        _, ref_l = O.generate(sd_cpu, cfg, inp["input_ids"][r:r + 1], inp["pixel_values"][r:r + 1], inp["attention_mask"][r:r + 1],
                              steps, return_logits=True, forced_tokens=toks[r:r + 1])
        probs = torch.softmax(ref_l[0].double() / temp, -1)  # [steps, V]
        for p_, name in ((top_p, "strict"), (top_p + 0.03, "loose")):
            srt, idx = O.top_p_filter(probs, p_)
            keep = torch.zeros_like(srt, dtype=torch.bool).scatter_(1, idx, srt > 0)
            inside = keep[torch.arange(steps), toks[r]]
            if name == "strict":
                strict += int(inside.sum())
                kept_sizes += keep.sum(-1).tolist()
            else:
                loose += int(inside.sum())
        total += steps
    print(f"[parity-3B-R2-b64] top-p draws inside the oracle kept set: strict {strict}/{total}, at top_p + 0.03 {loose}/{total}; "
          f"kept-set sizes {min(kept_sizes)}..{max(kept_sizes)}")
    _record("3b_224_R2_b64_top_p", {"draws": total, "in_oracle_kept_set": strict, "in_oracle_kept_set_top_p_plus_0.03": loose,
                                    "kept_set_size_min": min(kept_sizes), "kept_set_size_max": max(kept_sizes), "rows": list(ROWS),
                                    "workload": "3B-224 R2, B=64 sampling job (temp 0.8, top-p 0.9), 8 steps"})
    assert loose == total and strict >= 0.9 * total


# ---------------------------------------------------------------------------------------------------------------------
# 448 px / 896 px geometries at full widths (BASELINE configs[3] / [4])
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("image_size", [448, 896])
def test_3b_long_image_sequences_prefill_and_decode(setup_3b, image_size):
    """1024 / 4096 image tokens through the full-width SigLIP tower (multi-tile dh = 72 attention), projector, merge, Gemma
    prefill over S = 1028 / 4100 (many key tiles at dh = 256) and 4 decode steps over a 17 / 65-page KV cache.  Weights = the
    R1 state dict of the 224 model with a position-embedding table of the larger grid."""
    cfg224, _, sd_cpu224 = setup_3b
    cfg = paligemma_3b_config(image_size)
    n = (image_size // 14) ** 2
    g = torch.Generator().manual_seed(image_size)
    pos = (torch.randn(n, cfg["vision_config"]["hidden_size"], generator=g)).bfloat16().float()
    sd_cpu = dict(sd_cpu224)
    sd_cpu["vision_tower.model.embeddings.positional_embeddings.weight"] = pos
    model = build_model(cfg, {k: v.to(device="cuda", dtype=torch.bfloat16) for k, v in sd_cpu.items()})
    steps = 5
    inp = make_inputs(cfg, batch=1, prompt_len=4, seed=image_size)
    ref_t, ref_l = O.generate(sd_cpu, cfg, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], steps, return_logits=True)
    toks, logits = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), steps,
                                  return_logits=True, forced_tokens=ref_t)
    rows = []
    for t in range(steps):
        s = stats(logits[0, t], ref_l[0, t])
        mg = float(top2_margin(ref_l[0, t]))
        rows.append(dict(s, margin=mg))
        print(f"[parity-3B-{image_size}] step {t}: max_abs={s['max_abs']:.4g} ({100 * s['rel']:.3f}% of absmax {s['absmax']:.4g}) "
              f"cos={s['cos']:.6f} oracle margin {mg:.4g}")
        assert s["cos"] >= 0.9999 and s["rel"] <= 0.01, (t, s)
        if mg > 4 * s["max_abs"]:
            assert int(toks[0, t]) == int(ref_t[0, t])
    free = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), steps)
    _record(f"3b_{image_size}_R1_b1_teacher_forced", dict(_summ(rows), greedy_tokens_identical=free[0].tolist() == ref_t[0].tolist(),
                                                          workload=f"3B-{image_size} R1, B=1, S={inp['input_ids'].shape[1]}, prefill + 4 teacher-forced decode steps"))
    del model
    torch.cuda.empty_cache()
