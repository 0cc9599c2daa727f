"""End-to-end parity of the sm_100a model against the CPU fp32 oracle and the committed reference golden vectors.

Tolerances (bf16 operands, fp32 accumulation / residual stream / softmax; SURVEY.md 8(c) protocol):
  logits:  cosine >= 0.9999 and max-abs <= 1 % of the oracle's logit absmax in the well-conditioned regimes (R0, R1);
           cosine >= 0.999 / 3 % in the diffuse regime R2 (tiny model, where bf16 noise is amplified);
  greedy:  token sequences identical in R0/R1 (16 tokens tiny), and wherever the oracle's top-2 margin exceeds 4x the
           measured max-abs error under teacher forcing.
"""
import os

import numpy as np
import pytest
import torch

from parity_utils import ROOT, build_model, stats, top2_margin

pytestmark = pytest.mark.gpu

from oracle import paligemma_oracle as O  # noqa: E402
from paligemma_multimodal_system_b200.random_init import TINY_CONFIG, make_inputs, make_state_dict  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "tiny_reference.npz"))
TOL = {"R0": (0.9999, 0.01), "R1": (0.9999, 0.01), "R2": (0.999, 0.03)}


def _check(got, ref, regime, what):
    s = stats(got, ref)
    print(f"[parity] {what} {regime}: max_abs={s['max_abs']:.4g} ({100 * s['rel']:.3f}% of absmax {s['absmax']:.4g}) cos={s['cos']:.6f}")
    cos_min, rel_max = TOL[regime]
    assert s["cos"] >= cos_min and s["rel"] <= rel_max, (what, regime, s)
    return s


@pytest.mark.parametrize("regime", ["R0", "R1", "R2"])
def test_vision_tower_and_projector(regime):
    sd = make_state_dict(TINY_CONFIG, regime, seed=11)
    model = build_model(TINY_CONFIG, sd)
    inp = make_inputs(TINY_CONFIG, batch=2, prompt_len=6, seed=7)
    feats = model.vision_tower(inp["pixel_values"].cuda())
    assert feats.shape == (2, 256, 256) and feats.dtype == torch.float32
    _check(feats, O.siglip_forward(sd, TINY_CONFIG["vision_config"], inp["pixel_values"]), regime, "siglip")
    _check(feats[:, :4, :32], torch.from_numpy(G[f"{regime}_b2_vision_slice"]), regime, "siglip vs reference golden")
    proj = model.image_features(inp["pixel_values"].cuda())
    _check(proj, O.image_features(sd, TINY_CONFIG, inp["pixel_values"]), regime, "projector")


@pytest.mark.parametrize("regime", ["R0", "R1", "R2"])
def test_prefill_logits_all_positions_and_padding(regime):
    from paligemma_multimodal_system_b200.modeling_gemma import KVCache
    sd = make_state_dict(TINY_CONFIG, regime, seed=11)
    model = build_model(TINY_CONFIG, sd)
    inp = make_inputs(TINY_CONFIG, batch=2, prompt_len=6, seed=7)
    kv = KVCache()
    out = model(input_ids=inp["input_ids"].cuda(), pixel_values=inp["pixel_values"].cuda(),
                attention_mask=inp["attention_mask"].cuda(), kv_cache=kv)
    assert out["logits"].shape == (2, 262, 1281) and out["logits"].dtype == torch.float32 and out["kv_cache"] is kv
    assert kv.num_items() == int(G[f"{regime}_b2_num_items"]) == 262
    _check(out["logits"][:, [0, 255, -1], :], torch.from_numpy(G[f"{regime}_b2_logits_pos"]), regime, "prefill logits vs reference golden")
    ref = O.forward(sd, TINY_CONFIG, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], [])
    _check(out["logits"], ref, regime, "prefill logits (all positions)")
    # cached (post-RoPE) keys of layer 1, reference layout [B, Hkv, S, dh]
    k1 = kv.k_cache[1]
    assert k1.shape == (2, 1, 262, 64)
    _check(k1[:, 0, -3:, :], torch.from_numpy(G[f"{regime}_b2_kcache_l1_slice"]), regime, "k_cache slice")
    # padded prompt: pad embedding zeroed, position 1, never masked in attention (reference behaviour)
    ids, mask = inp["input_ids"].clone(), inp["attention_mask"].clone()
    ids[1, -1] = 0
    mask[1, -1] = 0
    out = model(input_ids=ids.cuda(), pixel_values=inp["pixel_values"].cuda(), attention_mask=mask.cuda(), kv_cache=None)
    assert "kv_cache" not in out
    _check(out["logits"][:, -1, :], torch.from_numpy(G[f"{regime}_b2_padded_logits_last"]), regime, "padded prompt")


@pytest.mark.parametrize("regime", ["R0", "R1", "R2"])
def test_reference_loop_through_forward_api(regime):
    """inference.py:45-79 driven through forward()/KVCache exactly as the reference loop does (B = 1, greedy)."""
    from paligemma_multimodal_system_b200.modeling_gemma import KVCache
    sd = make_state_dict(TINY_CONFIG, regime, seed=11)
    model = build_model(TINY_CONFIG, sd)
    inp = make_inputs(TINY_CONFIG, batch=1, prompt_len=4, seed=5)
    ids, mask, px = inp["input_ids"].cuda(), inp["attention_mask"].cuda(), inp["pixel_values"].cuda()
    gold_t, gold_l = G[f"{regime}_greedy_tokens"], torch.from_numpy(G[f"{regime}_step_logits"])
    kv = KVCache()
    worst = 0.0
    toks = []
    for step in range(16):
        out = model(input_ids=ids, pixel_values=px, attention_mask=mask, kv_cache=kv)
        kv = out["kv_cache"]
        logits = out["logits"][:, -1, :]
        s = stats(logits[0], gold_l[step])
        worst = max(worst, s["max_abs"])
        cos_min, rel_max = TOL[regime]
        assert s["cos"] >= cos_min and s["rel"] <= rel_max, (step, s)
        nxt = torch.argmax(logits, dim=-1, keepdim=True)
        toks.append(int(nxt))
        if float(top2_margin(gold_l[step])) > 4 * s["max_abs"]:
            assert int(nxt) == int(gold_t[step]), f"step {step}: argmax differs although the oracle margin is wide"
        # teacher forcing with the reference's own token keeps the comparison meaningful in every regime
        ids = torch.tensor([[int(gold_t[step])]], device="cuda")
        mask = torch.cat([mask, torch.ones((1, 1), device="cuda", dtype=mask.dtype)], dim=-1)
        assert kv.num_items() == 260 + step
    margins = top2_margin(gold_l)
    print(f"[parity] loop {regime}: worst max_abs {worst:.4g}; oracle top-2 margin min {margins.min():.4g} median {margins.median():.4g}; "
          f"tokens {toks} vs reference {gold_t.tolist()}")
    if regime in ("R0", "R1"):
        assert toks == gold_t.tolist()


@pytest.mark.parametrize("regime", ["R0", "R1"])
def test_generate_greedy_identical_to_reference(regime):
    sd = make_state_dict(TINY_CONFIG, regime, seed=11)
    model = build_model(TINY_CONFIG, sd)
    inp = make_inputs(TINY_CONFIG, batch=1, prompt_len=4, seed=5)
    for graph in (False, True):
        toks = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 16, use_cuda_graph=graph)
        assert toks.shape == (1, 16)
        assert toks[0].tolist() == G[f"{regime}_greedy_tokens"].tolist(), f"free-running greedy (graph={graph})"
    # second call reuses the captured graph and the cached KV storage
    toks2 = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 16)
    assert toks2[0].tolist() == G[f"{regime}_greedy_tokens"].tolist()


def test_graphed_prefill_of_the_latency_path_follows_its_inputs():
    """From the second generate() call of a geometry the whole prefill replays as one CUDA graph over static input buffers:
    different requests through the same graph must give what the eager path gives, errors included."""
    sd = make_state_dict(TINY_CONFIG, "R1", seed=11)
    model = build_model(TINY_CONFIG, sd)
    a = make_inputs(TINY_CONFIG, batch=2, prompt_len=4, seed=5)
    b = make_inputs(TINY_CONFIG, batch=2, prompt_len=4, seed=77)
    b["input_ids"][:, -1] = torch.tensor([300, 417])  # R1 echoes the last prompt token: make the answers differ
    run = lambda inp, **kw: model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 8, **kw)
    eager_a, eager_b = run(a, use_cuda_graph=False), run(b, use_cuda_graph=False)
    assert eager_a[0].tolist() == G["R1_greedy_tokens"][:8].tolist() and eager_a.tolist() != eager_b.tolist()
    first = run(a)            # eager warm-up of the graphed path
    second = run(b)           # capture + replay
    third = run(a)            # replay with other inputs
    assert model._graphs and any(st.get("prefill", {}).get("graph") is not None for st in model._graphs.values())
    assert first.tolist() == eager_a.tolist() and second.tolist() == eager_b.tolist() and third.tolist() == eager_a.tolist()
    _, lg = run(b, return_logits=True)
    _, le = run(b, return_logits=True, use_cuda_graph=False)
    _, le2 = run(b, return_logits=True, use_cuda_graph=False)
    noise, diff = stats(le2[:, 0], le[:, 0])["rel"], stats(lg[:, 0], le[:, 0])["rel"]
    print(f"[graphed prefill] prefill logits graph vs eager {diff:.2e} of absmax; eager vs eager (split-K order noise) {noise:.2e}")
    assert diff <= 1e-2  # same kernels: only the fp32 red.add order differs, within the stated logits tolerance
    bad = {k: v.clone() for k, v in a.items()}
    bad["input_ids"][1, 3] = 5  # 255 image tokens in row 1
    with pytest.raises(ValueError):
        run(bad)
    assert run(a).tolist() == eager_a.tolist()  # the graph is still usable after a rejected request


@pytest.mark.parametrize("image_size,batch", [(448, 2), (896, 1)])
def test_long_image_sequences_1024_and_4096_image_tokens(image_size, batch):
    """BASELINE configs[3] / [4] geometry (1024 / 4096 image tokens, position table, merge, many key tiles in both prefill
    attentions, a KV cache of 17 / 65 pages) on the tiny widths, so that the CPU oracle stays cheap: prefill + teacher-forced
    decode logits and free-running greedy tokens against the oracle."""
    import copy
    cfg = copy.deepcopy(TINY_CONFIG)
    cfg["vision_config"]["image_size"] = image_size
    sd = make_state_dict(cfg, "R1", seed=13)
    model = build_model(cfg, sd)
    inp = make_inputs(cfg, batch=batch, prompt_len=5, seed=3)
    inp["input_ids"][:, -1] = torch.arange(batch) + 200  # R1 echoes the last prompt token: rows differ
    T = 6
    ref_t, ref_l = O.generate(sd, cfg, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], T, return_logits=True)
    toks, logits = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), T,
                                  return_logits=True, forced_tokens=ref_t)
    for r in range(batch):
        _check(logits[r], ref_l[r], "R1", f"{image_size}px teacher-forced logits row {r}")
    free = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), T)
    assert free.cpu().tolist() == ref_t.tolist()
    assert model._graphs[next(iter(model._graphs))]["kv"].page_table.shape[1] >= (image_size // 14) ** 2 // 64 + 1


def test_batched_generate_rows_equal_single_row_runs():
    """B > 1 (not runnable in the reference): every row must reproduce its own B = 1 run, oracle as the judge."""
    sd = make_state_dict(TINY_CONFIG, "R2", seed=3)
    model = build_model(TINY_CONFIG, sd)
    inp = make_inputs(TINY_CONFIG, batch=5, prompt_len=5, seed=9)
    ref_t, ref_l = O.generate(sd, TINY_CONFIG, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 8, return_logits=True)
    toks, logits = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 8,
                                  return_logits=True, forced_tokens=ref_t)
    for r in range(5):
        _check(logits[r], ref_l[r], "R2", f"teacher-forced logits row {r}")
    t1, l1 = model.generate(inp["input_ids"][2:3].cuda(), inp["pixel_values"][2:3].cuda(), inp["attention_mask"][2:3].cuda(), 8,
                            return_logits=True, forced_tokens=ref_t[2:3])
    s = stats(l1[0], logits[2])
    # same kernels, different batch size: only the fp32 split-K summation order differs -- but in this deliberately
    # diffuse tiny regime an ulp-level difference occasionally flips a bf16 rounding of an activation, which shows up at
    # the 1 % level in the logits (both runs are within the oracle tolerance checked above)
    assert s["rel"] < 3e-2, s


def test_sampled_generate_stays_in_reference_kept_set():
    sd = make_state_dict(TINY_CONFIG, "R1", seed=3)
    model = build_model(TINY_CONFIG, sd)
    # bitwise reproducible decode (no split-K): the fp32 red.add order of the split-K GEMMs changes the logits in the last
    # bits from launch to launch, and an inverse-CDF draw sits within that distance of one of its V boundaries about once
    # per thousand draws, which would make the a == b check below flaky
    model.language_model.deterministic_decode = True
    inp = make_inputs(TINY_CONFIG, batch=4, prompt_len=5, seed=9)
    a = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 12, do_sample=True, seed=7)
    b = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 12, do_sample=True, seed=7)
    c = model.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 12, do_sample=True, seed=8)
    assert torch.equal(a, b) and not torch.equal(a, c)
    # each sampled token lies in the oracle's top-p kept set for the (teacher-forced) oracle distribution
    ref_t, ref_l = O.generate(sd, TINY_CONFIG, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 12,
                              return_logits=True, forced_tokens=a.cpu())
    probs = torch.softmax(ref_l / 0.8, -1)
    srt, idx = O.top_p_filter(probs.view(-1, probs.shape[-1]), 0.9)
    keep = torch.zeros_like(srt, dtype=torch.bool).scatter_(1, idx, srt > 0).view(4, 12, -1)
    # allow the (rare) boundary token whose probability differs within bf16 logit noise
    inside = keep[torch.arange(4)[:, None], torch.arange(12)[None, :], a.cpu()]
    assert inside.float().mean() >= 0.95, inside


def test_kvcache_reference_api():
    from paligemma_multimodal_system_b200.modeling_gemma import KVCache
    kv = KVCache()
    assert kv.num_items() == 0
    kv.allocate(2, 3, 1, 64, 100)
    k = torch.randn(2, 1, 70, 64, device="cuda").bfloat16().float()
    v = torch.randn(2, 1, 70, 64, device="cuda").bfloat16().float()
    K, V = kv.update(k, v, 0)
    assert kv.num_items() == 70 and torch.equal(K.float(), k) and torch.equal(V.float(), v)
    k2 = torch.randn(2, 1, 1, 64, device="cuda").bfloat16().float()
    K, V = kv.update(k2, k2, 0)  # grows like torch.cat(dim=-2) (modeling_gemma.py:54-55)
    assert kv.num_items() == 71 and torch.equal(K.float(), torch.cat([k, k2], -2))
    kv.update(torch.randn(2, 1, 200, 64, device="cuda"), torch.randn(2, 1, 200, 64, device="cuda"), 1)  # capacity growth
    assert torch.equal(kv.k_cache[0].float(), torch.cat([k, k2], -2)) and kv.k_cache[1].shape == (2, 1, 200, 64)


def test_input_validation():
    sd = make_state_dict(TINY_CONFIG, "R1", seed=11)
    model = build_model(TINY_CONFIG, sd)
    inp = make_inputs(TINY_CONFIG, batch=1, prompt_len=4, seed=5)
    ids = inp["input_ids"].clone()
    ids[0, 0] = 5  # 255 image tokens instead of 256: the reference silently mis-scatters; we refuse
    with pytest.raises(ValueError):
        model(input_ids=ids.cuda(), pixel_values=inp["pixel_values"].cuda(), attention_mask=inp["attention_mask"].cuda(), kv_cache=None)
    with pytest.raises(ValueError):
        model.vision_tower(torch.zeros(1, 3, 112, 112))
    with pytest.raises(RuntimeError):
        model.language_model.model.layers[0].mlp(torch.zeros(1))


@pytest.mark.parametrize("regime", ["R0", "R1"])
def test_prefill_with_rope_and_kv_append_in_the_qkv_epilogue(regime):
    """The prefill whose q/k/v projection rotates and appends in its own epilogue (pg_gemm_qkv_rope; automatic once the grid
    fills the GPU, forced here) against the oracle, against the path with the separate RoPE / append launch, and through decode
    steps that read the cache pages it wrote."""
    from paligemma_multimodal_system_b200.modeling_gemma import KVCache
    sd = make_state_dict(TINY_CONFIG, regime, seed=11)
    model = build_model(TINY_CONFIG, sd)
    inp = make_inputs(TINY_CONFIG, batch=3, prompt_len=6, seed=7)
    ids, px, mask = inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda()
    ref = O.forward(sd, TINY_CONFIG, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], [])
    outs, caches = {}, {}
    for fused in (False, True):
        model.language_model.fused_qkv_rope = fused
        kv = KVCache()
        outs[fused] = model(input_ids=ids, pixel_values=px, attention_mask=mask, kv_cache=kv)["logits"]
        caches[fused] = (kv.k_cache, kv.v_cache)
        _check(outs[fused], ref, regime, f"prefill logits, fused_qkv_rope={fused}")
    for l in range(TINY_CONFIG["text_config"]["num_hidden_layers"]):
        ka, kb = caches[False][0][l].float(), caches[True][0][l].float()
        assert (ka - kb).abs().max().item() <= 2 ** -6 * ka.abs().max().item()  # one bf16 ulp: libm vs hardware sin / cos
        va, vb = caches[False][1][l].float(), caches[True][1][l].float()
        assert (va - vb).abs().max().item() <= 2 ** -6 * va.abs().max().item()
    ref_t, ref_l = O.generate(sd, TINY_CONFIG, inp["input_ids"], inp["pixel_values"], inp["attention_mask"], 8, return_logits=True)
    model.language_model.fused_qkv_rope = True
    toks, logits = model.generate(ids, px, mask, 8, return_logits=True, forced_tokens=ref_t)
    for r in range(3):
        _check(logits[r], ref_l[r], regime, f"decode over pages written by the fused prefill, row {r}")
    assert model.generate(ids, px, mask, 8).cpu().tolist() == ref_t.tolist()
    model.language_model.fused_qkv_rope = None


@pytest.mark.parametrize("late", [False, True])
def test_decode_l2_prefetch_branch_changes_nothing(late, monkeypatch):
    """The forked L2 weight prefetch of decode_layers (pg_prefetch_l2 on a side stream / graph branch) is a scheduling aid:
    tokens and logits of a generate() job are bitwise identical with and without it, eagerly and through the CUDA graphs,
    for both fork points (after the q/k/v launch; after the attention launch for large KV caches)."""
    from paligemma_multimodal_system_b200.modeling_gemma import GemmaForCausalLM, KVCache
    sd = make_state_dict(TINY_CONFIG, "R1", seed=5)
    inp = make_inputs(TINY_CONFIG, batch=3, prompt_len=6, seed=2)
    args = (inp["input_ids"].cuda(), inp["pixel_values"].cuda(), inp["attention_mask"].cuda(), 20)
    if late:  # make every cache look large to the policy
        monkeypatch.setattr(KVCache, "capacity", property(lambda self: 0 if self.page_table is None else 1 << 20))
    outs = []
    for pf_bytes in (0, 1 << 20):
        monkeypatch.setattr(GemmaForCausalLM, "l2_prefetch_bytes", pf_bytes)
        model = build_model(TINY_CONFIG, sd)
        model.language_model.deterministic_decode = True  # no split-K: bitwise reproducible logits
        for graph in (False, True):
            toks = model.generate(*args, do_sample=False, use_cuda_graph=graph)
            toks2, logits = model.generate(*args, do_sample=False, use_cuda_graph=False, return_logits=True)
            outs.append((toks.clone(), toks2.clone(), logits.clone()))
    for o in outs[1:]:
        assert torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) and torch.equal(outs[0][2], o[2])
