"""GPU parity of the layer stack behind `DiffusionModel`, p_sample and the denoising loop.

Tolerances: p_sample and the schedule are bit-exact (same f32 operation order as the
reference); the SIMT stack matches the f64 oracle to f32 noise; the tcgen05 stack (bf16
operands per linear) matches within 1e-2 * sqrt(n_linears) relative Frobenius error — one
bf16-operand rounding per linear, accumulated over the depth of the stack."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F = np.float32


@pytest.fixture(scope="module")
def ctx():
    import dllm_b200
    c = dllm_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def O():
    from oracle import pyoracle
    return pyoracle


def beq(a, b):
    """bit-for-bit f32 equality; NaNs must sit at the same places (their payload bits are not
    part of the reference's semantics: x86 and the GPU produce different default NaNs)."""
    a, b = np.asarray(a, F), np.asarray(b, F)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb]))


def build_stack(ctx, O, rng, dims, bits=4, group=128, with_bias=True):
    """dims: [d0, d1, ..., dn] -> n linears; weights N(0, 1/K) to keep unit gain through the stack."""
    from dllm_b200 import QWeight
    gpu_layers, ref_layers = [], []
    for K, N in zip(dims[:-1], dims[1:]):
        w = (rng.standard_normal((K, N)) / np.sqrt(K)).astype(F)
        b = (rng.standard_normal(N) * 0.1).astype(F) if with_bias else None
        gpu_layers.append(QWeight.quantize(ctx, w, bits, group, b))
        c, s, z = O.quantize_weight_grouped(w, bits, group)
        ref_layers.append((c, s, z, b, group))
    return gpu_layers, ref_layers


def ref_forward64(O, x, ref_layers):
    h = np.asarray(x, np.float64)
    for (c, s, z, b, group) in ref_layers:
        wd = O.dequantize_weight_grouped(c, s, z, group).astype(np.float64)
        h = h @ wd + (b.astype(np.float64) if b is not None else 0.0)
    return h


def test_p_sample_bit_exact(ctx, O):
    from dllm_b200.diffuse_llm import DiffuseLLM, DiffusionConfig, QuantizedDiffusionModel, BetaSchedule
    rng = np.random.default_rng(0)
    layers, _ = build_stack(ctx, O, rng, [128, 128])
    for sched in (BetaSchedule.Linear, BetaSchedule.Quadratic, BetaSchedule.Cosine):
        cfg = DiffusionConfig(num_timesteps=100, hidden_size=128, beta_schedule=sched)
        model = QuantizedDiffusionModel(layers, 128, cfg, ctx)
        llm = DiffuseLLM(cfg, ctx)
        betas = O.beta_schedule(sched.value, 100)
        assert np.array_equal(cfg.create_beta_schedule().view(np.uint32), betas.view(np.uint32))
        x = rng.standard_normal((6, 256)).astype(F)
        pred = rng.standard_normal((6, 256)).astype(F)
        z = rng.standard_normal((6, 256)).astype(F)
        for t in ([7] * 6, [99] * 6, [1] * 6, [3, 9, 27, 81, 99, 250]):
            out = llm.p_sample(model, x, t, pred, z)
            assert beq(out, O.p_sample(x, pred, z, t, betas, True))
        # t[0] == 0: no noise for the whole batch; guarded rows keep x_t; unguarded is the literal NaN
        assert beq(llm.p_sample(model, x, [0] * 6, pred, z), O.p_sample(x, pred, z, [0] * 6, betas, True))
        assert beq(llm.p_sample(model, x, [0] * 6, pred, z), x)
        lit = llm.p_sample(model, x, [0, 5, 5, 5, 5, 5], pred, z, guard_t0=False)
        exp = O.p_sample(x, pred, z, [0, 5, 5, 5, 5, 5], betas, False)
        assert np.array_equal(np.isnan(lit), np.isnan(exp)) and beq(lit[1:], exp[1:])
        model.close()


def test_add_noise_bit_exact(ctx, O):
    """DiffuseLLM::add_noise (lib.rs:1100-1137) on the GPU == the oracle, bit for bit, for the three schedules,
    per-row timesteps (clamped past T-1, :1123), and the device-pointer form."""
    from dllm_b200.diffuse_llm import DiffuseLLM, DiffusionConfig, QuantizedDiffusionModel, BetaSchedule
    rng = np.random.default_rng(5)
    layers, _ = build_stack(ctx, O, rng, [128, 128])
    for sched in (BetaSchedule.Linear, BetaSchedule.Quadratic, BetaSchedule.Cosine):
        cfg = DiffusionConfig(num_timesteps=100, hidden_size=128, beta_schedule=sched)
        model = QuantizedDiffusionModel(layers, 128, cfg, ctx)
        llm = DiffuseLLM(cfg, ctx)
        betas = O.beta_schedule(sched.value, 100)
        x = rng.standard_normal((6, 384)).astype(F)
        nz = rng.standard_normal((6, 384)).astype(F)
        x[0, :4] = [np.nan, np.inf, -np.inf, 0.0]
        for t in ([0] * 6, [99] * 6, [3, 9, 27, 81, 99, 250], [0, 1, 2, 3, 4, 5]):
            noisy, noise_out = llm.add_noise(model, x, t, nz)
            assert noise_out is not None and beq(noise_out, nz)
            assert beq(noisy, O.add_noise(x, t, nz, betas))
        with pytest.raises(Exception):
            llm.add_noise(model, x, [1] * 6, None)
        # device-pointer form, one timestep for the batch
        n = x.size
        dx, dn, do = ctx.malloc(n * 4), ctx.malloc(n * 4), ctx.malloc(n * 4)
        ctx.h2d(dx, x); ctx.h2d(dn, nz)
        ctx._ck(ctx._lib.dllm_add_noise_dev(ctx.h, model.h, dx, dn, 42, 6, 384, do))
        assert beq(ctx.d2h(do, x.shape, F), O.add_noise(x, [42] * 6, nz, betas))
        for p in (dx, dn, do):
            ctx.free(p)
        model.close()


def test_noise_generator_bit_exact(ctx, O):
    """"dllm_noise v1" (csrc/noise.cuh): the device generator and the oracle's C restatement agree bit for bit — every
    step is exact or one correctly rounded f32 operation in a fixed order, no libm / SFU function — for odd / even
    offsets, odd lengths and several seeds / streams."""
    for seed, stream, n, first in ((42, 0, 100001, 0), (42, 7, 4096, 1), (0, 2 ** 40 + 3, 33, 1_000_001), (2 ** 63 + 5, 999, 7, 6)):
        g = ctx.noise_fill(seed, stream, n, first)
        assert beq(g, O.noise_normal(seed, stream, n, first)), (seed, stream, n, first)
    z = ctx.noise_fill(42, 1, 1 << 20)
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.std()) - 1.0) < 5e-3 and np.all(np.isfinite(z))
    assert ctx.noise_fill(1, 2, 0).size == 0


@pytest.mark.parametrize("use_graph", [False, True])
def test_sample_seeded_matches_oracle(ctx, O, use_graph):
    """The seeded loop (noise generated in the p_sample kernel, step replayed from a CUDA graph) == the oracle's loop
    (lib.rs:853-927) fed with the same generator's streams: timestep t = stream t, initial x = stream num_steps."""
    from dllm_b200 import PATH_SIMT
    from dllm_b200.diffuse_llm import DiffuseLLM, DiffusionConfig, QuantizedDiffusionModel
    rng = np.random.default_rng(3)
    H, seq, batch, steps, seed = 128, 3, 5, 6, 42
    layers, ref = build_stack(ctx, O, rng, [H, 256, H])
    cfg = DiffusionConfig(num_timesteps=50, hidden_size=H, use_kv_cache=False)
    model = QuantizedDiffusionModel(layers, H, cfg, ctx, PATH_SIMT)
    llm = DiffuseLLM(cfg, ctx)
    n = batch * seq * H
    x0 = O.noise_normal(seed, steps, n).reshape(batch, -1)
    noises = [O.noise_normal(seed, t, n).reshape(batch, -1) for t in range(steps)]
    exp = O.sample(x0, ref, H, steps, O.beta_schedule(O.BETA_LINEAR, 50), noises, True)
    r0 = ctx.graph_replays
    out = llm.sample_seeded(model, (batch, seq), steps, seed, use_graph=use_graph)
    assert np.linalg.norm(out - exp) <= 1e-4 * np.linalg.norm(exp)
    assert ctx.graph_replays - r0 == (steps - 1 if use_graph else 0)
    # graph replay and eager launches are the same kernels: bit-identical; a second run reuses the captured graph
    again = llm.sample_seeded(model, (batch, seq), steps, seed, use_graph=not use_graph)
    assert beq(out, again)
    # an explicit x0 overrides the generated one
    x1 = rng.standard_normal((batch, seq * H)).astype(F)
    o1 = llm.sample_seeded(model, (batch, seq), steps, seed, x0=x1, use_graph=use_graph)
    e1 = O.sample(x1, ref, H, steps, O.beta_schedule(O.BETA_LINEAR, 50), noises, True)
    assert np.linalg.norm(o1 - e1) <= 1e-4 * np.linalg.norm(e1)
    model.close()


@pytest.mark.parametrize("bits", [4, 8])
def test_stack_forward_simt(ctx, O, bits):
    from dllm_b200 import PATH_SIMT
    from dllm_b200.diffuse_llm import QuantizedDiffusionModel
    rng = np.random.default_rng(bits)
    dims = [256, 256, 512, 256]
    layers, ref = build_stack(ctx, O, rng, dims, bits)
    model = QuantizedDiffusionModel(layers, 256, ctx=ctx, path=PATH_SIMT)
    x = rng.standard_normal((3, 256 * 4)).astype(F)          # batch 3, seq 4
    y = model.forward(x, np.zeros(3))
    y64 = ref_forward64(O, x.reshape(-1, 256), ref).reshape(3, -1)
    assert y.shape == x.shape
    assert np.linalg.norm(y - y64) <= 1e-5 * np.linalg.norm(y64)
    # the reference composition itself (dequantize_tensor ∘ x.dot(W)+b in f32)
    y32 = O.model_forward(x.reshape(-1, 256), ref).reshape(3, -1)
    assert np.allclose(y, y32, rtol=0, atol=2e-5 * np.abs(y64).max())
    assert beq(model.forward_with_cache(x, None, None, None), y)     # lib.rs:815-824
    model.close()


def test_stack_forward_umma(ctx, O):
    from dllm_b200 import PATH_UMMA
    from dllm_b200.diffuse_llm import QuantizedDiffusionModel
    rng = np.random.default_rng(11)
    dims = [256, 256, 256, 512, 256]
    layers, ref = build_stack(ctx, O, rng, dims, 4)
    model = QuantizedDiffusionModel(layers, 256, ctx=ctx, path=PATH_UMMA)
    x = rng.standard_normal((8, 256 * 32)).astype(F)         # 256 tokens
    y = model.forward(x)
    y64 = ref_forward64(O, x.reshape(-1, 256), ref).reshape(8, -1)
    rel = np.linalg.norm(y - y64) / np.linalg.norm(y64)
    assert rel <= 1e-2 * np.sqrt(len(layers)), rel
    model.close()


def test_simple_diffusion_model_is_the_reference_layer(ctx, O):
    """SimpleDiffusionModel: one linear x·W+b, weights N(0,1)*0.02, bias 0 (lib.rs:775-813)."""
    from dllm_b200 import PATH_SIMT
    from dllm_b200.diffuse_llm import SimpleDiffusionModel
    m = SimpleDiffusionModel(256, 256, bits=8, seed=42, ctx=ctx, path=PATH_SIMT)
    assert m.weights.shape == (256, 256) and np.all(m.bias == 0)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 256)).astype(F)
    y = m.forward(x, np.zeros(4))
    c, s, z = O.quantize_weight_grouped(m.weights, 8, 128)
    y32 = O.qlinear_forward(x, c, s, z, m.bias, 128)
    assert np.allclose(y, y32, rtol=0, atol=1e-5)
    # 8-bit quantization error against the unquantized reference layer stays small
    assert np.abs(y - (x @ m.weights)).max() < 2e-2
    m.close()


@pytest.mark.parametrize("path_name", ["SIMT", "UMMA"])
def test_sample_loop_matches_oracle(ctx, O, path_name):
    """DiffuseLLM::sample without cache (lib.rs:853-927), injected noise, 8 steps."""
    import dllm_b200
    from dllm_b200.diffuse_llm import DiffuseLLM, DiffusionConfig, QuantizedDiffusionModel
    path = getattr(dllm_b200, "PATH_" + path_name)
    rng = np.random.default_rng(5)
    hidden, seq, batch, steps = 128, 16, 8, 8
    layers, ref = build_stack(ctx, O, rng, [hidden, 256, hidden], 4)
    cfg = DiffusionConfig(num_timesteps=50, hidden_size=hidden, use_kv_cache=False)
    model = QuantizedDiffusionModel(layers, hidden, cfg, ctx, path)
    llm = DiffuseLLM(cfg, ctx)
    x0 = rng.standard_normal((batch, hidden * seq)).astype(F)
    noises = rng.standard_normal((steps, batch, hidden * seq)).astype(F)
    out = llm.sample(model, (batch, seq), steps, None, x0=x0, noises=noises)
    betas = O.beta_schedule(O.BETA_LINEAR, 50)
    exp = O.sample(x0, ref, hidden, steps, betas, noises, True)
    assert out.shape == (batch, hidden * seq) and np.all(np.isfinite(out))
    rel = np.linalg.norm(out - exp) / np.linalg.norm(exp)
    # the loop feeds each step's output back in: per-step error compounds over `steps`
    tol = 1e-4 if path_name == "SIMT" else 1e-2 * np.sqrt(2 * steps)
    assert rel <= tol, rel
    model.close()


def test_sample_with_phase_aware_cache(ctx, O):
    """The cached branch of sample (lib.rs:885-921): progressive decode bits, cache re-quantized per step."""
    from dllm_b200 import PATH_SIMT
    from dllm_b200.diffuse_llm import DiffuseLLM, DiffusionConfig, QuantizedDiffusionModel, KVCacheEntry
    rng = np.random.default_rng(9)
    hidden = 128
    layers, ref = build_stack(ctx, O, rng, [hidden, hidden], 4)
    cfg = DiffusionConfig(num_timesteps=50, hidden_size=hidden, num_layers=2)
    model = QuantizedDiffusionModel(layers, hidden, cfg, ctx, PATH_SIMT)
    llm = DiffuseLLM(cfg, ctx)
    k = rng.standard_normal((2, 4, hidden)).astype(F)
    llm.kv_cache["c"] = KVCacheEntry(k, k.copy(), 8, 4, ctx)
    x0 = rng.standard_normal((2, hidden * 2)).astype(F)
    noises = rng.standard_normal((6, 2, hidden * 2)).astype(F)
    out = llm.sample(model, (2, 2), 6, "c", x0=x0, noises=noises)
    exp = O.sample(x0, ref, hidden, 6, O.beta_schedule(O.BETA_LINEAR, 50), noises, True)
    assert np.linalg.norm(out - exp) <= 1e-4 * np.linalg.norm(exp)       # the simple model ignores the cache
    cache = llm.kv_cache["c"]
    assert not cache.is_prefill_phase and cache.decode_quant_bits == O.progressive_bits(6, 0)[0]
    model.close()


def test_cpp_host_mirror_harness():
    """The C++ mirror of the reference interface (host/dllm.hpp) run as the reference's own unit tests."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "diffusion-llm-rs_b200", "host", "host_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "HOST_TEST_OK" in out.stdout, out.stdout + out.stderr
