#!/usr/bin/env python
"""Headline benchmark: PaliGemma-3B-pt-224 decode tokens/s (+ prefill ms/image) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N ...            # the reference's own CPU path on the host cores

Workload = BASELINE.json configs[2]: 3B-224 architecture (random init, bf16-representable, regime R2), 64 requests per
GPU (weak scaling: batch-sharded replicas, no collective on the data path), prompt `<image>*256 <bos> t1 t2 \\n`
(S = 260), temperature 0.8 / top-p 0.9 sampling, 128 generated tokens.  One "step" = one whole generate() job
(prefill of 64 images + 127 graph-replayed decode steps).

  value       decode tokens/s over all ranks: (K * B * 127 * N) / max-over-ranks(sum of device-timed decode segments),
              inputs resident in HBM.  `prefill_ms_per_image` is the same for the prefill segments.
  e2e         tokens/s through the public API from pinned HOST buffers: H2D of pixels/ids/mask and D2H of the tokens
              inside the timed region, prefill included (all 128 tokens counted).
  roofline    dominant kernel (gate||up weight-streaming tcgen05 GEMM of the decode step) timed alone with CUDA events, rotating over the 18 layers' weights
              (2.4 GB > L2) against the measured HBM peak; `roofline_step` is the whole decode step (5.40 GB algorithmic
              bytes) from the graph replays of the timed region.
  configs     the other BASELINE.json configurations measured in the same process: cfg2 (B = 1 greedy-32 latency),
              cfg4 (448 px, 32 requests), cfg5 (896 px, 8 requests), and the strong-scaling split of configs[2]
              (global batch 64 => 64 / N requests per GPU).
  parity      measured GPU-vs-oracle figures at the 3B widths (profiles/parity_r02.json, written by tests/test_bench_3b_gpu.py).
  cpu_baseline  the UNMODIFIED reference (staged under baseline/_ref, kind "reference") or, when that copy is absent, the
              oracle port (kind "port") on a bounded sample, rank 0, N = 1.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 64
PROMPT_LEN = 4
NEW_TOKENS = 128
TEMPERATURE, TOP_P = 0.8, 0.9
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
DOMINANT_KERNEL = "gemm_tcgen05_kernel<64, 1, 1, 0>"  # swap-AB (weights on the UMMA M axis), 64 token columns, 8 epilogue warps
TRAFFIC_CAPTURE = os.path.join(ROOT, "profiles", "r02i_decode_gemm_ncu_full.csv")


def algorithmic_bytes_per_decode_step(cfg, batch, kv_len_avg):
    """every LM weight once (bf16) + KV read + KV write; vision tower / dead logits excluded."""
    tc = cfg["text_config"]
    D, F, L, V = tc["hidden_size"], tc["intermediate_size"], tc["num_hidden_layers"], tc["vocab_size"]
    Hq, Hkv, dh = tc["num_attention_heads"], tc["num_key_value_heads"], tc["head_dim"]
    per_layer = (Hq + 2 * Hkv) * dh * D + D * D + 3 * D * F + 2 * D
    weights = 2 * (L * per_layer + D + V * D + V)
    kv_read = batch * L * 2 * kv_len_avg * Hkv * dh * 2
    kv_write = batch * L * 2 * Hkv * dh * 2
    return weights + kv_read + kv_write


def algorithmic_flops_per_image(cfg, S):
    """full non-causal attention as in the reference, last-position lm_head only."""
    vc, tc = cfg["vision_config"], cfg["text_config"]
    Dv, Fv, Lv = vc["hidden_size"], vc["intermediate_size"], vc["num_hidden_layers"]
    N = (vc["image_size"] // vc["patch_size"]) ** 2
    D, F, L, V = tc["hidden_size"], tc["intermediate_size"], tc["num_hidden_layers"], tc["vocab_size"]
    Hq, Hkv, dh = tc["num_attention_heads"], tc["num_key_value_heads"], tc["head_dim"]
    siglip = N * (2 * 3 * vc["patch_size"] ** 2 * Dv + Lv * (2 * (4 * Dv * Dv + 2 * Dv * Fv) + 4 * N * Dv))
    proj = N * 2 * Dv * D
    gemma = S * L * (2 * (Hq * dh * D + 2 * Hkv * dh * D + D * D + 3 * D * F) + 4 * S * Hq * dh)
    head = 2 * D * V
    return float(siglip + proj + gemma + head)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("bf16_tflops_sustained", 1400.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1400.0


def ncu_capture_of_dominant_kernel(path=TRAFFIC_CAPTURE, kernel=DOMINANT_KERNEL, grid="(256, 1, 1)"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed `ncu --set full`
    capture, plus that capture's (isolated, cold, serialised) duration.  The row must NAME the kernel `roofline.kernel`
    names, at the gate||up grid: a capture of any other kernel yields no traffic figure."""
    import csv
    try:
        with open(path) as f:
            rows = list(csv.reader(f))
    except OSError:
        return None
    hdr = rows[0]
    for row in rows[1:]:
        rec = dict(zip(hdr, row))
        if kernel not in rec.get("Kernel Name", "") or rec.get("Grid Size", "").replace(" ", "") != grid.replace(" ", ""):
            continue
        tot, dur = 0.0, None
        for name, val in rec.items():
            if name.startswith(("dram__bytes_read.sum", "dram__bytes_write.sum")):
                unit = name[name.index("[") + 1:name.index("]")].lower()
                tot += float(val) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
            if name.startswith("gpu__time_duration.sum"):
                unit = name[name.index("[") + 1:name.index("]")].lower()
                dur = float(val) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1.0)
        return {"traffic": tot, "isolated_us": dur, "file": os.path.relpath(path, ROOT)}
    return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx, pw = [], set(), None, 0.0
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw = max(pw, float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "power_w_max": pw, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_gpu_model(cfg, seed=0):
    import copy
    from paligemma_multimodal_system_b200.modeling_paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration
    from paligemma_multimodal_system_b200.random_init import make_state_dict
    sd = make_state_dict(cfg, "R2", seed=seed, device="cuda", dtype=torch.bfloat16)
    model = PaliGemmaForConditionalGeneration(PaliGemmaConfig(**copy.deepcopy(cfg)), device="meta").eval()
    model.load_state_dict(sd, strict=True, assign=True)
    model.tie_weights()
    model.pack()
    return model, sd


def time_dominant_kernel(model, batch, iters=3):
    """gate||up decode GEMM exactly as the decode step launches it, alone, CUDA events on the launch stream, cold L2 (consecutive
    launches read different layers' 134 MB weight matrices)."""
    from paligemma_multimodal_system_b200 import _lib
    lm = model.language_model
    pk = lm._packed
    c = lm.text_config
    x = (torch.randn(batch, c.hidden_size, device="cuda") * 0.1).bfloat16()
    out = torch.empty(batch, c.intermediate_size, device="cuda", dtype=torch.bfloat16)

    def launch(lw):
        _lib.gemm(x, lw["gu_w"], out, mode=_lib.EPI_GEGLU, swap=1)

    for lw in pk["layers"]:
        launch(lw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 0
    for _ in range(iters):
        for lw in pk["layers"]:
            launch(lw)
            n += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    bytes_alg = 2 * c.intermediate_size * c.hidden_size * 2 + batch * c.hidden_size * 2 + batch * c.intermediate_size * 2
    return ms, bytes_alg


# ----------------------------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (baseline/_ref, staged by __graft_entry__.build() where /root/reference exists) or the
# oracle port
# ----------------------------------------------------------------------------------------------------------------------
def reference_staged():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("modeling_paligemma.py", "modeling_gemma.py", "modeling_siglip.py", "inference.py"))


def load_reference(cfg, sd_cpu):
    """The reference's own classes from baseline/_ref (an unmodified copy of /root/reference/*.py), its stock constructor,
    load_state_dict and tie_weights.  `fire` (its CLI dependency, absent from this image) is stubbed: the CLI is not on the path."""
    import copy
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.modules.setdefault("fire", types.SimpleNamespace(Fire=lambda f: None))
    import inference as ref_inference
    from modeling_gemma import KVCache
    from modeling_paligemma import PaliGemmaConfig, PaliGemmaForConditionalGeneration
    model = PaliGemmaForConditionalGeneration(PaliGemmaConfig(**copy.deepcopy(cfg))).eval()
    model.load_state_dict(sd_cpu, strict=True, assign=True)
    model.tie_weights()
    return model, KVCache, ref_inference


@torch.no_grad()
def cpu_reference_sample(ref, inp, new_tokens, do_sample):
    """The stock loop of inference.py:45-79 at model level (B = 1: the reference cannot decode a batch), its own
    forward / KVCache / _sample_top_p, stdout silenced (the reference prints shapes).  Vision tower re-run every step, as it does."""
    model, KVCache, ref_inference = ref
    ids, mask, px = inp["input_ids"], inp["attention_mask"], inp["pixel_values"]
    kv = KVCache()
    t = [time.perf_counter()]
    for _ in range(new_tokens + 1):
        with contextlib.redirect_stdout(io.StringIO()):
            out = model(input_ids=ids, pixel_values=px, attention_mask=mask, kv_cache=kv)
        kv = out["kv_cache"]
        logits = out["logits"][:, -1, :]
        if do_sample:
            nxt = ref_inference._sample_top_p(torch.softmax(logits / TEMPERATURE, dim=-1), TOP_P)
        else:
            nxt = torch.argmax(logits, dim=-1, keepdim=True)
        ids = nxt.squeeze(0).unsqueeze(-1)
        mask = torch.cat([mask, torch.ones((1, 1), dtype=mask.dtype)], dim=-1)
        t.append(time.perf_counter())
    return dict(prefill_s_per_image=t[1] - t[0], decode_tok_s=new_tokens / (t[-1] - t[1]), seconds=t[-1] - t[0])


def cpu_oracle_sample(cfg, sd_cpu, rows, new_tokens, threads):
    """Times the oracle port (CPU fp32) on `rows` requests: prefill seconds/image and decode tokens/s."""
    from oracle import paligemma_oracle as O
    from paligemma_multimodal_system_b200.random_init import make_inputs
    torch.set_num_threads(threads)
    inp = make_inputs(cfg, batch=rows, prompt_len=PROMPT_LEN, seed=1)
    t0 = time.perf_counter()
    kv = []
    feats = O.image_features(sd_cpu, cfg, inp["pixel_values"])
    logits = O.forward(sd_cpu, cfg, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], kv, last_only=True, image_feats=feats)
    t1 = time.perf_counter()
    ids, mask = logits[:, -1].argmax(-1, keepdim=True), inp["attention_mask"]
    for _ in range(new_tokens):
        mask = torch.cat([mask, torch.ones((rows, 1), dtype=mask.dtype)], -1)
        logits = O.forward(sd_cpu, cfg, ids, inp["pixel_values"], mask, kv, last_only=True, image_feats=feats)
        ids = logits[:, -1].argmax(-1, keepdim=True)
    t2 = time.perf_counter()
    return dict(prefill_s_per_image=(t1 - t0) / rows, decode_tok_s=rows * new_tokens / (t2 - t1), seconds=t2 - t0)


def cpu_baseline_entry(cfg, sd_cpu, threads, tokens, do_sample):
    """One bounded CPU sample for the `cpu_baseline` key: the staged reference when present, else the oracle port."""
    from paligemma_multimodal_system_b200.random_init import make_inputs
    torch.set_num_threads(threads)
    if reference_staged():
        ref = load_reference(cfg, sd_cpu)
        inp = make_inputs(cfg, batch=1, prompt_len=PROMPT_LEN, seed=1)
        r = cpu_reference_sample(ref, inp, tokens, do_sample)
        kind = "reference"
        sample = (f"1 request (the reference asserts B = 1), prefill + {tokens} {'top-p' if do_sample else 'greedy'} decode tokens through the "
                  "UNMODIFIED reference (baseline/_ref: stock forward / KVCache loop, vision tower re-run per step as it does), fp32")
    else:
        r = cpu_oracle_sample(cfg, sd_cpu, 1, tokens, threads)
        kind = "port"
        sample = f"1 request, prefill + {tokens} greedy decode tokens, fp32 oracle port (baseline/_ref not staged), vision tower not re-run"
    return {"value": r["decode_tok_s"], "unit": "tokens/s", "cores": threads, "kind": kind,
            "prefill_ms_per_image": 1e3 * r["prefill_s_per_image"], "sample": sample}


def measure_serving(model, cfg, n_requests=256, slots=64, stage=32, min_admit=16, steps_per_replay=8, greedy=False, static=True):
    """a ragged request stream (text prompts of 2..12 tokens, token budgets uniform in [16, 128] standing
    in for EOS-terminated answers) through serving.ContinuousBatcher, beside static batching of the same stream through
    generate() (arrival order, `slots` at a time, each batch runs to its longest budget; one prompt length per batch, i.e.
    with perfect length bucketing).  Useful tokens = sum of the budgets in both arms; wall clock around the whole stream,
    host scheduling, H2D of the pixels and D2H of the tokens included; the first pass of each arm is the warm-up."""
    from paligemma_multimodal_system_b200.random_init import make_inputs, make_requests
    from paligemma_multimodal_system_b200.serving import ContinuousBatcher
    reqs = make_requests(cfg, n_requests, 2, 12, seed=3)
    budgets = torch.randint(16, 129, (n_requests,), generator=torch.Generator().manual_seed(5)).tolist()
    useful = sum(budgets)
    gen = dict(do_sample=not greedy, temperature=TEMPERATURE, top_p=TOP_P, seed=1)
    N = (cfg["vision_config"]["image_size"] // cfg["vision_config"]["patch_size"]) ** 2
    res = {"requests": n_requests, "slots": slots, "stage": stage, "min_admit": min_admit, "useful_tokens": useful,
           "stream": "prompts of N+2..N+12 tokens, budgets uniform 16..128, " + ("greedy" if greedy else "top-p 0.9 temp 0.8")}
    if static:
        inp = make_inputs(cfg, batch=slots, prompt_len=7, seed=9)
        dev = {k: v.cuda() for k, v in inp.items()}
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for lo in range(0, n_requests, slots):
                model.generate(dev["input_ids"], dev["pixel_values"], dev["attention_mask"], max(budgets[lo:lo + slots]), **gen).cpu()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        res["static_useful_tokens_per_s"] = useful / dt
    cb = ContinuousBatcher(model, num_slots=slots, max_prompt_len=N + 12, max_new_tokens=128, min_admit=min_admit, stage=stage,
                           steps_per_replay=steps_per_replay, **gen)
    for _ in range(2):
        cb.reset_stats()
        for (ids, px), m in zip(reqs, budgets):
            cb.submit(ids, px, m)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = cb.run()
        dt = time.perf_counter() - t0
    assert sum(len(v) for v in out.values()) == useful
    st = cb.stats
    res.update({"useful_tokens_per_s": useful / dt, "requests_per_s": n_requests / dt, "prefill_groups": st["prefill_groups"],
                "decode_steps": st["decode_steps"], "slot_occupancy": useful / max(1, st["decode_steps"] * slots),
                "host_ms": {k[2:-2]: round(1e3 * v, 1) for k, v in st.items() if k.startswith("t_")}})
    return res


def time_generate(model, cfg, batch, new_tokens, gen, runs, warmup, seed=100):
    """Device-timed generate() jobs with HBM-resident inputs: summed prefill / decode milliseconds over `runs` jobs."""
    from paligemma_multimodal_system_b200.random_init import make_inputs
    inp = make_inputs(cfg, batch=batch, prompt_len=PROMPT_LEN, seed=seed)
    dev = {k: v.cuda() for k, v in inp.items()}
    for _ in range(warmup):
        model.generate(dev["input_ids"], dev["pixel_values"], dev["attention_mask"], new_tokens, **gen)
    torch.cuda.synchronize()
    pre = dec = 0.0
    for _ in range(runs):
        tm = {}
        model.generate(dev["input_ids"], dev["pixel_values"], dev["attention_mask"], new_tokens, timings=tm, **gen)
        pre += tm["prefill_ms"]
        dec += tm["decode_ms"]
    return pre, dec, inp["input_ids"].shape[1]


def run_reference_arm(args, cfg):
    """`--impl reference`: the reference's own CPU implementation of the path on the host cores, same config / metric, each
    step a bounded sample: one request (the reference cannot batch its decode), prefill + a few top-p decode tokens through the
    stock forward / KVCache loop.  Falls back to the oracle port (kind "port") only when baseline/_ref was not staged."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from paligemma_multimodal_system_b200.random_init import make_inputs, make_state_dict
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = make_state_dict(cfg, "R2", seed=0, device="cpu")
    staged = reference_staged()
    toks = 16 if args.steps <= 4 else (8 if args.steps <= 10 else 4)  # >= 16 timed decode tokens in total, minutes in total
    if staged:
        ref = load_reference(cfg, sd)
        inp = make_inputs(cfg, batch=1, prompt_len=PROMPT_LEN, seed=1)
        run = lambda n: cpu_reference_sample(ref, inp, n, True)
        kind = "reference"
        sample = (f"per step: 1 request (reference asserts B = 1), prefill + {toks} top-p decode tokens through the UNMODIFIED reference "
                  "(baseline/_ref), stock forward / KVCache / _sample_top_p, vision tower re-run per step as the reference does; 3B-224 fp32")
    else:
        run = lambda n: cpu_oracle_sample(cfg, sd, 1, n, threads)
        kind = "port"
        sample = f"per step: 1 request, prefill + {toks} greedy decode tokens, oracle port (baseline/_ref not staged); 3B-224 fp32"
    for _ in range(min(args.warmup, 1)):
        run(1)
    t0 = time.perf_counter()
    res = [run(toks) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    dec = toks * len(res) / sum(toks / r["decode_tok_s"] for r in res)
    pre = sum(r["prefill_s_per_image"] for r in res) / len(res)
    line = {"impl": "reference", "metric": "decode_tokens_per_s", "value": dec, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "prefill_ms_per_image": 1e3 * pre,
            "config": {"workload": "PaliGemma-3B-pt-224 random-init (R2), top-p 0.9 temp 0.8 (BASELINE configs[2]); CPU sample", "sample": sample},
            "cpu_baseline": {"value": dec, "unit": "tokens/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": dec, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if staged:  # the port beside it, labelled (B = 1, vision tower not re-run: an upper bound on what the reference code could do)
        r = cpu_oracle_sample(cfg, sd, 1, 4, threads)
        line["cpu_port"] = {"value": r["decode_tok_s"], "unit": "tokens/s", "cores": threads, "kind": "port",
                            "sample": "1 request, prefill + 4 greedy tokens, oracle port, vision tower not re-run"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="requests per GPU")
    ap.add_argument("--new-tokens", type=int, default=NEW_TOKENS)
    ap.add_argument("--image-size", type=int, default=224)
    ap.add_argument("--greedy", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-serving", action="store_true", help="skip the continuous-batching sample (N = 1 only)")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg2 / cfg4 / cfg5 / strong-scaling measurements")
    args = ap.parse_args()

    from paligemma_multimodal_system_b200.random_init import make_inputs, paligemma_3b_config
    cfg = paligemma_3b_config(args.image_size)
    if args.impl == "reference":
        return run_reference_arm(args, cfg)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from paligemma_multimodal_system_b200 import _lib
    _lib.require_device()

    B, T = args.batch, args.new_tokens
    model, sd = build_gpu_model(cfg, seed=0)
    # rank r serves rows [r*B, (r+1)*B) of the global batch (B per GPU: weak scaling): independent requests, no exchange
    from paligemma_multimodal_system_b200.sharding import shard_bounds
    lo, hi = shard_bounds(B * world, world, rank)
    assert hi - lo == B
    inp = make_inputs(cfg, batch=B, prompt_len=PROMPT_LEN, seed=100 + rank)
    S = inp["input_ids"].shape[1]
    host = {k: v.pin_memory() for k, v in inp.items()}
    dev = {k: v.cuda() for k, v in inp.items()}
    gen = dict(do_sample=not args.greedy, temperature=TEMPERATURE, top_p=TOP_P, seed=1234)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm ----------------
    for _ in range(args.warmup):
        model.generate(dev["input_ids"], dev["pixel_values"], dev["attention_mask"], T, **gen)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.lib().pg_launch_count()
    pre_ms = dec_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        tm = {}
        model.generate(dev["input_ids"], dev["pixel_values"], dev["attention_mask"], T, timings=tm, **gen)
        pre_ms += tm["prefill_ms"]
        dec_ms += tm["decode_ms"]
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    eager_launches = _lib.lib().pg_launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    # ---------------- end-to-end arm (host buffers, copies inside the timed region) ----------------
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    d2h = 0
    for _ in range(args.steps):
        ids = host["input_ids"].cuda(non_blocking=True)
        msk = host["attention_mask"].cuda(non_blocking=True)
        px = host["pixel_values"].cuda(non_blocking=True)
        toks = model.generate(ids, px, msk, T, **gen).cpu()
        d2h = toks.numel() * toks.element_size()
    e3.record()
    barrier()
    e2e_ms = e2.elapsed_time(e3)
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    # ---------------- steady-state decode (reported beside the in-job number, not instead of it) ----------------
    # Inside generate() the first ~50 decode steps follow the 1 kW prefill burst and run at reduced SM clocks (the power
    # governor needs ~80 ms to come back); replaying the captured 8-step graph on its own shows the kernel-limited rate.
    steady_ms = 0.0
    stt = next((v for v in model._graphs.values() if v.get("graph_k") is not None), None)
    if stt is not None:
        kvc = stt["kv"]
        n_rep = max(1, (T - 1) // 8)
        for _ in range(2):  # first pass: warm-up at steady clocks
            kvc.counters[0].fill_(S + 1); kvc.counters[1].fill_(S); kvc.counters[2].fill_(S + 1); stt["step"].zero_()
            torch.cuda.synchronize()
            e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e4.record()
            for _ in range(n_rep):
                stt["graph_k"].replay()
            e5.record()
            torch.cuda.synchronize()
            steady_ms = e4.elapsed_time(e5) / (8 * n_rep)

    # ---------------- dominant kernel alone, and the continuous-batching sample (rank 0; both need the 224-px model) ----------------
    k_ms = k_bytes = None
    serving = None
    if rank == 0:
        k_ms, k_bytes = time_dominant_kernel(model, B)
        if world == 1 and not args.no_serving and args.image_size == 224:
            try:
                serving = measure_serving(model, cfg, slots=B, greedy=args.greedy)
            except Exception as e:  # a reported extra, never allowed to take the headline line down
                serving = {"error": f"{type(e).__name__}: {e}"}

    # ---------------- the other BASELINE configurations, same process ----------------
    # [cfg2 prefill, cfg2 decode, strong prefill, strong decode, cfg4 prefill, cfg4 decode, cfg5 prefill, cfg5 decode]
    extra = [0.0] * 8
    extra_meta = {}
    if not args.no_configs and args.image_size == 224 and B == BATCH_PER_GPU:
        runs = 2
        greedy = dict(do_sample=False)
        pre2, dec2, _ = time_generate(model, cfg, 1, 32, greedy, runs=5, warmup=3)            # cfg-2: latency path
        extra[0], extra[1] = pre2 / 5, dec2 / 5
        bs = max(1, BATCH_PER_GPU // world)                                                   # strong scaling of configs[2]
        if world == 1:
            extra[2], extra[3] = pre_ms / args.steps, dec_ms / args.steps
        else:
            p, d, _ = time_generate(model, cfg, bs, T, gen, runs=runs, warmup=2, seed=300 + rank)
            extra[2], extra[3] = p / runs, d / runs
        extra_meta["strong_batch_per_gpu"] = bs
        for slot, (px_size, bx) in enumerate(((448, 32), (896, 8))):                          # cfg-4 / cfg-5
            model._graphs.clear()
            model = None  # release the previous geometry's packed weights, graphs and KV caches
            torch.cuda.empty_cache()
            cfg_x = paligemma_3b_config(px_size)
            model, _ = build_gpu_model(cfg_x, seed=0)
            p, d, Sx = time_generate(model, cfg_x, bx, T, gen, runs=runs, warmup=2, seed=500 + rank)
            extra[4 + 2 * slot], extra[5 + 2 * slot] = p / runs, d / runs
            extra_meta[px_size] = dict(cfg=cfg_x, S=Sx, B=bx)
        barrier()

    times = torch.tensor([total_ms, pre_ms, dec_ms, e2e_ms, steady_ms] + extra, device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, pre_ms, dec_ms, e2e_ms, steady_ms, *extra = times.tolist()

    if rank == 0:
        K = args.steps
        dec_steps = T - 1
        decode_tok_s = K * B * dec_steps * world / (dec_ms / 1e3)
        hbm_peak, peak_src, tf_peak = peaks()
        step_bytes = algorithmic_bytes_per_decode_step(cfg, B, S + T / 2)
        step_ms = dec_ms / (K * dec_steps)
        step_gbs = step_bytes / (step_ms * 1e-3) / 1e9
        k_gbs = k_bytes / (k_ms * 1e-3) / 1e9
        cap = ncu_capture_of_dominant_kernel() if B == 64 else None
        # per decode step: embed + 7 per layer (norm, qkv, attention, o, norm, gate||up, down) + final norm + head + sampler + advance
        #   + one pg_prefetch_l2 launch per layer on the forked branch (L2 weight prefetch)
        pf_on = model.language_model.l2_prefetch_bytes >= 16
        graph_kernels = (8 if pf_on else 7) * cfg["text_config"]["num_hidden_layers"] + 5
        flops_img = algorithmic_flops_per_image(cfg, S)
        line = {
            "metric": "decode_tokens_per_s", "value": decode_tok_s, "unit": "tokens/s", "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "prefill_ms_per_image": pre_ms / (K * B), "decode_ms_per_token_step": step_ms,
            "config": {"workload": f"PaliGemma-3B-pt-{args.image_size} random-init (R2), {B} requests/GPU, S={S}, "
                                   f"{'greedy' if args.greedy else 'top-p 0.9 temp 0.8'}, {T} new tokens"
                                   + (" (BASELINE configs[2])" if args.image_size == 224 and B == 64 and T == 128 and not args.greedy else ""),
                       "global_batch": B * world, "parallelism": f"dp{world} replicas, batch-sharded, no collective",
                       "l2_policy": "inputs larger than L2 (5.0 GB of weights streamed per decode step)"},
            "e2e": {"value": K * B * T * world / (e2e_ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "generated tokens / wall incl. H2D, prefill, decode, D2H"},
            "gpu_launches": int(eager_launches + K * max(T - 1, 0) * graph_kernels),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": DOMINANT_KERNEL + " (swap-AB gate||up decode GEMM, GeGLU epilogue)",
                         "achieved": k_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": k_gbs / hbm_peak,
                         "traffic": cap["traffic"] if cap else None, "peak_source": peak_src,
                         "us_per_launch": 1e3 * k_ms, "algorithmic_bytes": k_bytes,
                         "timing": "CUDA events around 54 back-to-back launches on the launch stream (programmatic dependent launch: the "
                                   "next launch's weight prefetch overlaps the running one), cold L2 (2.4 GB of weights in rotation)",
                         "ncu_isolated_us": cap["isolated_us"] if cap else None, "traffic_source": cap["file"] if cap else None},
            "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": step_gbs / hbm_peak,
                              "algorithmic_bytes": step_bytes, "kernels_per_layer": 7,
                              "l2_prefetch_branch": bool(pf_on)},
            "roofline_prefill": {"bound": "tensor", "achieved": flops_img / (pre_ms / (K * B) * 1e-3) / 1e12,
                                 "peak": tf_peak, "unit": "TFLOP/s",
                                 "frac": flops_img / (pre_ms / (K * B) * 1e-3) / 1e12 / tf_peak,
                                 "algorithmic_flops_per_image": flops_img,
                                 "peak_source": peak_src + " bf16_tflops_sustained (cuBLAS, seconds-long loop)"},
            "steady_state_decode": None if steady_ms <= 0 else {
                "ms_per_token_step": steady_ms, "tokens_per_s": B * world / (steady_ms / 1e3),
                "frac_of_hbm_peak": step_bytes / (steady_ms * 1e-3) / 1e9 / hbm_peak,
                "note": "8-step decode graph replayed back to back without the preceding prefill burst (SM clocks at max)"},
        }
        if extra_meta:
            bs = extra_meta["strong_batch_per_gpu"]
            sb = algorithmic_bytes_per_decode_step(cfg, bs, S + T / 2)
            configs = {
                "cfg2_latency": {"workload": "3B-224, batch 1, greedy 32 tokens (BASELINE configs[1]); per GPU, max over ranks",
                                 "prefill_ms": extra[0], "ms_per_token": extra[1] / 31, "tokens_32_ms": extra[0] + extra[1],
                                 "frac": algorithmic_bytes_per_decode_step(cfg, 1, S + 16) / (extra[1] / 31 * 1e-3) / 1e9 / hbm_peak},
                "strong_scaling": {"workload": f"BASELINE configs[2] with the GLOBAL batch fixed at 64: {bs} requests/GPU on {world} GPU(s)",
                                   "prefill_ms_per_image": extra[2] / bs, "ms_per_token_step": extra[3] / dec_steps,
                                   "decode_tokens_per_s": bs * world * dec_steps / (extra[3] / 1e3),
                                   "roofline_step_frac": sb / (extra[3] / dec_steps * 1e-3) / 1e9 / hbm_peak},
            }
            for px_size, name, ref_cfg in ((448, "cfg4_448", "configs[3]"), (896, "cfg5_896", "configs[4]")):
                m = extra_meta[px_size]
                p_ms, d_ms = extra[4 if px_size == 448 else 6], extra[5 if px_size == 448 else 7]
                fl = algorithmic_flops_per_image(m["cfg"], m["S"])
                by = algorithmic_bytes_per_decode_step(m["cfg"], m["B"], m["S"] + T / 2)
                configs[name] = {"workload": f"3B-{px_size}, {m['B']} requests/GPU, S={m['S']}, top-p, {T} new tokens (BASELINE {ref_cfg})",
                                 "prefill_ms_per_image": p_ms / m["B"],
                                 "roofline_prefill": {"achieved": fl / (p_ms / m["B"] * 1e-3) / 1e12, "unit": "TFLOP/s",
                                                      "frac": fl / (p_ms / m["B"] * 1e-3) / 1e12 / tf_peak},
                                 "decode_tokens_per_s": m["B"] * world * dec_steps / (d_ms / 1e3),
                                 "roofline_step": {"achieved": by / (d_ms / dec_steps * 1e-3) / 1e9, "unit": "GB/s",
                                                   "frac": by / (d_ms / dec_steps * 1e-3) / 1e9 / hbm_peak}}
            line["configs"] = configs
        ppath = os.path.join(ROOT, "profiles", "parity_r02.json")
        if os.path.exists(ppath):
            try:
                line["parity"] = dict(json.load(open(ppath)), source="profiles/parity_r02.json (written by tests/test_bench_3b_gpu.py on a B200)")
            except ValueError:
                pass
        if serving is not None:
            line["serving"] = serving
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sd_cpu = {k: v.float().cpu() for k, v in sd.items()}
            try:
                line["cpu_baseline"] = cpu_baseline_entry(cfg, sd_cpu, threads, 8, do_sample=False)
            except Exception as e:
                line["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
