"""Mirror of diffuse_llm_rs::diffuse_llm (diffuse-llm-rs/src/lib.rs): the layer interface
(`DiffusionModel`), the quantized layer stack that implements it on the GPU, the phase-aware KV
cache entry, and the denoising loop (`DiffuseLLM::sample`, `p_sample`)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from enum import Enum
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from .quantization import QuantizedKVCacheEntry
from .runtime import Context, QWeight, default_context


class BetaSchedule(Enum):
    """lib.rs:109-117"""
    Linear = L.BETA_LINEAR
    Quadratic = L.BETA_QUADRATIC
    Cosine = L.BETA_COSINE


@dataclass
class QuantizationConfig:
    """phase-aware KV precision, lib.rs:85-105"""
    prefill_bits: int = 8
    decode_bits: int = 4
    progressive_precision: bool = True
    min_decode_bits: int = 2


@dataclass
class DiffusionConfig:
    """lib.rs:52-81, defaults :476-493"""
    num_timesteps: int = 1000
    hidden_size: int = 768
    num_layers: int = 12
    num_attention_heads: int = 12
    vocab_size: int = 50257
    max_sequence_length: int = 1024
    beta_start: float = 0.0001
    beta_end: float = 0.02
    beta_schedule: BetaSchedule = BetaSchedule.Linear
    use_kv_cache: bool = True
    kv_quant_bits: int = 4
    max_cache_size: int = 2 * 1024 * 1024 * 1024
    use_phase_aware_quant: bool = True
    quant_config: QuantizationConfig = field(default_factory=QuantizationConfig)

    def create_beta_schedule(self):
        """lib.rs:554-593 (f32, host arithmetic inside the library)"""
        betas = np.empty(self.num_timesteps, np.float32)
        L.check(L.lib().dllm_beta_schedule(self.beta_schedule.value, self.num_timesteps, self.beta_start,
                                           self.beta_end, betas.ctypes.data))
        return betas


class KVCacheEntry:
    """lib.rs:122-313: f32 K/V plus up to two quantized copies (prefill / decode bits)."""

    def __init__(self, keys, values, prefill_bits: int, decode_bits: int, ctx: Context | None = None):
        self._ctx = ctx or default_context()
        self.keys = np.ascontiguousarray(keys, np.float32)
        self.values = np.ascontiguousarray(values, np.float32)
        self.prefill_quant_bits, self.decode_quant_bits = prefill_bits, decode_bits
        self.is_prefill_phase = True
        self.seq_len = self.keys.shape[1]
        self.prefill_quantized = self._q(prefill_bits) if prefill_bits > 0 else None    # :145-153
        self.decode_quantized = self._q(decode_bits) if decode_bits > 0 else None       # :155-163

    def _q(self, bits):
        if self.keys.size == 0:
            return None
        return QuantizedKVCacheEntry(self.keys, self.values, bits, self._ctx)

    def _active(self):
        return self.prefill_quantized if self.is_prefill_phase else self.decode_quantized

    def get_keys(self):
        q = self._active()
        return q.dequantize_keys() if q is not None else self.keys.copy()

    def get_values(self):
        q = self._active()
        return q.dequantize_values() if q is not None else self.values.copy()

    def set_phase(self, is_prefill: bool):
        self.transition_phase(is_prefill)

    def get_current_quant_bits(self) -> int:
        return self.prefill_quant_bits if self.is_prefill_phase else self.decode_quant_bits

    def transition_phase(self, is_prefill: bool):
        if self.is_prefill_phase == is_prefill:
            return
        self.is_prefill_phase = is_prefill
        if not is_prefill and self.decode_quant_bits > 0 and self.decode_quantized is None:   # :228-235
            self.decode_quantized = self._q(self.decode_quant_bits)

    def update(self, new_keys, new_values):
        """:246-276 — re-quantizes at both precisions"""
        self.keys = np.ascontiguousarray(new_keys, np.float32)
        self.values = np.ascontiguousarray(new_values, np.float32)
        self.seq_len = self.keys.shape[1]
        if self.prefill_quant_bits > 0:
            self.prefill_quantized = self._q(self.prefill_quant_bits)
        if self.decode_quant_bits > 0:
            self.decode_quantized = self._q(self.decode_quant_bits)

    def memory_usage(self) -> int:
        total = 0
        for q in (self.prefill_quantized, self.decode_quantized):    # :283-296
            if q is not None:
                total += q.memory_usage()
        return total if total else self.keys.size * 4 + self.values.size * 4

    def len(self) -> int:
        return self.seq_len

    def is_empty(self) -> bool:
        return self.seq_len == 0


class DeviceKVCacheEntry:
    """KVCacheEntry (lib.rs:122-313) resident in HBM (`dllm_kvcache_*`): the f32 keys / values plus the prefill- and the
    decode-precision quantized copies all live on the device.  `update_dev` / `append_dev` take device pointers, `get_dev`
    decodes the active phase's copy straight into the consumer's device buffers — the cached branch of the sampling loop
    (`DiffuseLLM.sample_cached_dev`) copies nothing over PCIe.  `scheme`: L.KV_TENSOR_B is the reference's per-tensor
    quantizer (quantization.rs:140-157: every update re-quantizes everything); L.KV_ROW_D / L.KV_FIXED_C entries can grow
    token by token (`append_dev` quantizes only the new tokens)."""

    def __init__(self, ctx: Context, layers: int, hidden: int, capacity: int, prefill_bits: int, decode_bits: int,
                 scheme: int = L.KV_TENSOR_B):
        self._ctx = ctx
        self.layers, self.hidden, self.capacity, self.scheme = layers, hidden, capacity, scheme
        h = C.c_void_p()
        ctx._ck(ctx._lib.dllm_kvcache_create(ctx.h, layers, hidden, capacity, prefill_bits, decode_bits, scheme, C.byref(h)))
        self.h = h

    # -- device-pointer interface --
    def update_dev(self, keys_dev: int, values_dev: int, seq: int):
        """:246-276.  Dense [layers, seq, hidden] f32 device tensors."""
        self._ctx._ck(self._ctx._lib.dllm_kvcache_update_dev(self._ctx.h, self.h, keys_dev, values_dev, seq))

    def refresh_dev(self):
        """update() with the entry's own tensors (what SimpleDiffusionModel::update_kv_cache hands back, :826-835)."""
        self._ctx._ck(self._ctx._lib.dllm_kvcache_update_dev(self._ctx.h, self.h, None, None, self.len()))

    def append_dev(self, keys_new_dev: int, values_new_dev: int, t_new: int):
        self._ctx._ck(self._ctx._lib.dllm_kvcache_append_dev(self._ctx.h, self.h, keys_new_dev, values_new_dev, t_new))

    def get_dev(self, keys_out_dev: Optional[int], values_out_dev: Optional[int]):
        """get_keys / get_values (:176-205) into dense [layers, seq, hidden] device buffers."""
        self._ctx._ck(self._ctx._lib.dllm_kvcache_get_dev(self._ctx.h, self.h, keys_out_dev, values_out_dev))

    def set_phase(self, is_prefill: bool):
        self._ctx._ck(self._ctx._lib.dllm_kvcache_set_phase(self._ctx.h, self.h, int(is_prefill)))

    transition_phase = set_phase

    def set_decode_bits(self, bits: int):
        """:899-903 — `decode_quant_bits = bits; decode_quantized = None`"""
        self._ctx._ck(self._ctx._lib.dllm_kvcache_set_decode_bits(self._ctx.h, self.h, bits))

    def _info(self):
        s, p, b, m = C.c_size_t(), C.c_int32(), C.c_uint8(), C.c_size_t()
        L.check(self._ctx._lib.dllm_kvcache_info(self.h, C.byref(s), C.byref(p), C.byref(b), C.byref(m)))
        return s.value, bool(p.value), b.value, m.value

    def len(self) -> int:
        return self._info()[0]

    def is_empty(self) -> bool:
        return self.len() == 0

    @property
    def is_prefill_phase(self) -> bool:
        return self._info()[1]

    def get_current_quant_bits(self) -> int:
        return self._info()[2]

    def memory_usage(self) -> int:
        return self._info()[3]

    # -- host conveniences (tests, interop with the numpy KVCacheEntry) --
    def _host(self, which: int):
        shape = (self.layers, self.len(), self.hidden)
        n = shape[0] * shape[1] * shape[2]
        if n == 0:
            return np.empty(shape, np.float32)
        with self._ctx.lock:
            d = self._ctx.malloc(n * 4)
            try:
                self.get_dev(d if which == 0 else None, d if which == 1 else None)
                out = self._ctx.d2h(d, shape, np.float32)
                self._ctx.sync()
            finally:
                self._ctx.free(d)
        return out

    def get_keys(self):
        return self._host(0)

    def get_values(self):
        return self._host(1)

    def update(self, new_keys, new_values):
        k = np.ascontiguousarray(new_keys, np.float32)
        v = np.ascontiguousarray(new_values, np.float32)
        n = k.size
        with self._ctx.lock:
            dk, dv = self._ctx.malloc(max(n, 1) * 4), self._ctx.malloc(max(n, 1) * 4)
            try:
                self._ctx.h2d(dk, k)
                self._ctx.h2d(dv, v)
                self.update_dev(dk, dv, k.shape[1])
                self._ctx.sync()
            finally:
                self._ctx.free(dk)
                self._ctx.free(dv)

    def export_copy(self, prefill: bool):
        """(key_codes, value_codes, key_scale, key_zp, value_scale, value_zp) of one quantized copy, one code per u8 — the
        reference's QuantizedTensor view — or None if that copy does not exist."""
        kv = self._ctx._lib.dllm_kvcache_copy(self.h, int(prefill))
        if not kv:
            return None
        rows = self.layers * self.len()
        n = rows * self.hidden
        kc, vc = np.empty(n, np.uint8), np.empty(n, np.uint8)
        m = rows if self.scheme == L.KV_ROW_D else 1
        ks, kz, vs, vz = (np.empty(m, np.float32) for _ in range(4))
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_kv_export(self._ctx.h, kv, kc.ctypes.data, vc.ctypes.data, ks.ctypes.data,
                                                        kz.ctypes.data, vs.ctypes.data, vz.ctypes.data))
        return kc, vc, ks, kz, vs, vz

    def close(self):
        if getattr(self, "h", None):
            self._ctx.sync()
            self._ctx._lib.dllm_kvcache_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DiffusionModel:
    """trait DiffusionModel: Send + Sync (lib.rs:748-772)"""

    def forward(self, x, t):
        raise NotImplementedError

    def forward_with_cache(self, x, t, keys, values):
        raise NotImplementedError

    def update_kv_cache(self, x, t, cache: KVCacheEntry):
        raise NotImplementedError


class QuantizedDiffusionModel(DiffusionModel):
    """A stack of quantized linears `x·W+b` (each one the reference's SimpleDiffusionModel op,
    lib.rs:806-813) resident in HBM; forward runs on the GPU.  x [batch, hidden*seq] is viewed as
    [batch*seq, hidden] tokens; the stack's output width equals its input width."""

    def __init__(self, layers: Sequence[QWeight], hidden: int, config: DiffusionConfig | None = None,
                 ctx: Context | None = None, path: int = L.PATH_AUTO):
        self._ctx = ctx or default_context()
        self.layers = list(layers)
        self.hidden = hidden
        self.config = config or DiffusionConfig(hidden_size=hidden)
        self.path = path
        arr = (C.c_void_p * len(self.layers))(*[w.h for w in self.layers])
        h = C.c_void_p()
        self._ctx._ck(self._ctx._lib.dllm_model_create(
            self._ctx.h, hidden, arr, len(self.layers), self.config.num_timesteps,
            self.config.beta_schedule.value, self.config.beta_start, self.config.beta_end, C.byref(h)))
        self.h = h

    @classmethod
    def from_f32(cls, weights: Sequence[np.ndarray], biases: Sequence[Optional[np.ndarray]], bits: int,
                 group: int = 128, **kw):
        ctx = kw.get("ctx") or default_context()
        layers = [QWeight.quantize(ctx, w, bits, group, b) for w, b in zip(weights, biases)]
        return cls(layers, weights[0].shape[0], **kw)

    def forward(self, x, t=None):
        x = np.ascontiguousarray(x, np.float32)
        batch, feat = x.shape
        out = np.empty_like(x)
        tt = np.ascontiguousarray(t if t is not None else np.zeros(batch), np.uint64)
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_model_forward(self._ctx.h, self.h, x.ctypes.data, tt.ctypes.data, batch,
                                                            feat, out.ctypes.data, self.path))
        return out

    def forward_dev(self, x_dev: int, batch: int, feat: int, out_dev: int):
        self._ctx._ck(self._ctx._lib.dllm_model_forward_dev(self._ctx.h, self.h, x_dev, batch, feat, out_dev, self.path))

    def forward_with_cache(self, x, t, keys, values):
        return self.forward(x, t)            # lib.rs:815-824: the cache is ignored

    def update_kv_cache(self, x, t, cache: KVCacheEntry):
        return cache.keys.copy(), cache.values.copy()   # lib.rs:826-835

    # -- the cached branch on device tensors (DiffuseLLM.sample_cached_dev) --
    def update_kv_cache_dev(self, x_dev: int, t: int, batch: int, feat: int, cache: "DeviceKVCacheEntry"):
        """lib.rs:826-835: the reference layer hands the cache's own tensors back -> ("same",).  A model that produces keys /
        values returns ("update", keys_dev, values_dev, seq) (dense [layers, seq, hidden]) or ("append", k_new_dev,
        v_new_dev, t_new)."""
        return ("same",)

    def forward_with_cache_dev(self, x_dev: int, t: int, batch: int, feat: int, keys_dev: int, values_dev: int, seq: int,
                               out_dev: int):
        self.forward_dev(x_dev, batch, feat, out_dev)     # lib.rs:815-824: the cache is ignored

    def denoise_step_dev(self, x_dev: int, z_dev: Optional[int], t: int, batch: int, feat: int, guard_t0=True):
        self._ctx._ck(self._ctx._lib.dllm_denoise_step_dev(self._ctx.h, self.h, x_dev, z_dev, t, batch, feat,
                                                           int(guard_t0), self.path))

    def close(self):
        if getattr(self, "h", None):
            self._ctx.sync()
            self._ctx._lib.dllm_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SimpleDiffusionModel(QuantizedDiffusionModel):
    """lib.rs:775-813: one linear layer, weights N(0,1)*0.02, zero bias — held quantized in HBM."""

    def __init__(self, input_dim: int, output_dim: int, bits: int = 4, group: int = 128, seed: int | None = None,
                 weights=None, bias=None, **kw):
        rng = np.random.default_rng(seed)
        self.weights = (np.asarray(weights, np.float32) if weights is not None
                        else (rng.standard_normal((input_dim, output_dim)) * 0.02).astype(np.float32))   # :792-796
        self.bias = np.asarray(bias, np.float32) if bias is not None else np.zeros(output_dim, np.float32)  # :798
        ctx = kw.get("ctx") or default_context()
        g = group if (group and input_dim % group == 0) else 0
        layer = QWeight.quantize(ctx, self.weights, bits, g, self.bias)
        super().__init__([layer], input_dim, **kw)


class DiffuseLLM:
    """The sampler half of lib.rs (`impl DiffuseLLM`, :853-955, :1100-1215)."""

    def __init__(self, config: DiffusionConfig | None = None, ctx: Context | None = None):
        self.config = config or DiffusionConfig()
        self._ctx = ctx or default_context()
        self.kv_cache = {}
        self.cache_memory_usage = 0          # AtomicUsize of lib.rs:847

    # -- p_sample: lib.rs:1152-1215, noise injected --
    def p_sample(self, model: QuantizedDiffusionModel, x_t, t, noise_pred, noise=None, guard_t0=True):
        x_t = np.ascontiguousarray(x_t, np.float32)
        noise_pred = np.ascontiguousarray(noise_pred, np.float32)
        batch, feat = x_t.shape
        tt = np.ascontiguousarray(t, np.uint64)
        z = np.ascontiguousarray(noise, np.float32) if noise is not None else None
        out = np.empty_like(x_t)
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_p_sample(self._ctx.h, model.h, x_t.ctypes.data, noise_pred.ctypes.data,
                                                       z.ctypes.data if z is not None else None, tt.ctypes.data,
                                                       batch, feat, int(guard_t0), out.ctypes.data))
        return out

    # -- add_noise: lib.rs:1100-1137, noise injected --
    def add_noise(self, model: QuantizedDiffusionModel, x_start, t, noise):
        """(noisy, noise) with noisy = x_start*sqrt(alpha_bar_t) + noise*sqrt(1 - alpha_bar_t).  The reference draws
        the noise from an unseeded thread_rng when it is None (:1107-1109); here it must be supplied."""
        x_start = np.ascontiguousarray(x_start, np.float32)
        if noise is None:
            raise L.InvalidParams(L.ERR_INVALID_PARAMS, "add_noise: supply the noise (the reference's own draw is unseeded)")
        noise = np.ascontiguousarray(noise, np.float32)
        batch, feat = x_start.shape
        tt = np.ascontiguousarray(t, np.uint64)
        out = np.empty_like(x_start)
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_add_noise(self._ctx.h, model.h, x_start.ctypes.data, noise.ctypes.data,
                                                        tt.ctypes.data, batch, feat, out.ctypes.data))
        return out, noise

    def sample_seeded(self, model: QuantizedDiffusionModel, shape, num_steps: Optional[int] = None, seed: int = 42,
                      x0=None, guard_t0: bool = True, use_graph: bool = True):
        """DiffuseLLM::sample without cache (lib.rs:853-927) with the noise drawn on the device from the counter-based
        generator "dllm_noise v1" (timestep t = stream t, initial x = stream num_steps): nothing is uploaded per step and
        the step is replayed from one CUDA graph.  42 is the one seed the reference uses (examples/diffusion_example.rs:69)."""
        batch, seq_len = shape
        num_steps = num_steps if num_steps is not None else self.config.num_timesteps
        feat = self.config.hidden_size * seq_len
        x = np.ascontiguousarray(x0, np.float32) if x0 is not None else None
        out = np.empty((batch, feat), np.float32)
        with self._ctx.lock:
            self._ctx._ck(self._ctx._lib.dllm_sample_seeded(self._ctx.h, model.h, x.ctypes.data if x is not None else None,
                                                            seed, batch, feat, num_steps, int(guard_t0), model.path,
                                                            int(use_graph), out.ctypes.data))
        return out

    # -- cache management: lib.rs:977-1084 (host-side policy over the entries; the entries' tensors stay where they are) --
    def get_or_init_cache(self, cache_id: str, batch_size: int = 1):
        """:977-986"""
        if cache_id not in self.kv_cache:
            self.kv_cache[cache_id] = self.init_kv_cache(batch_size)
        return self.kv_cache[cache_id]

    def update_kv_cache(self, cache_id: str, keys, values):
        """:988-1043.  Accounting exactly as written there: the incoming entry is charged keys.len()*4*2 bytes against
        max_cache_size BEFORE it is quantized, eviction frees at least the overshoot, an existing entry is updated in place
        (usage grows by max(0, f32 size - its previous quantized size)), a new entry is charged its quantized size.
        (The reference's `KVCacheEntry::new(keys, values, kv_quant_bits)` at :1030 lacks an argument and does not compile;
        the evident intent — the same width for both phases — is what runs here.)"""
        if not self.config.use_kv_cache:
            return
        keys = np.ascontiguousarray(keys, np.float32)
        values = np.ascontiguousarray(values, np.float32)
        entry_size = keys.size * 4 * 2
        new_usage = self.cache_memory_usage + entry_size
        if new_usage > self.config.max_cache_size:
            self.evict_oldest_entries(new_usage - self.config.max_cache_size)
        entry = self.kv_cache.get(cache_id)
        if entry is not None:
            old_size = entry.memory_usage()
            entry.update(keys, values)
            self.cache_memory_usage += max(0, entry_size - old_size)              # saturating_sub, :1021-1024
        else:
            entry = self._new_entry(keys, values, self.config.kv_quant_bits, self.config.kv_quant_bits)
            self.kv_cache[cache_id] = entry
            self.cache_memory_usage += entry.memory_usage()

    def _new_entry(self, keys, values, prefill_bits, decode_bits):
        return KVCacheEntry(keys, values, prefill_bits, decode_bits, self._ctx)

    def evict_oldest_entries(self, bytes_to_free: int):
        """:1046-1073 — despite its name the reference evicts the LARGEST entries first (sort by memory usage, descending)
        until at least bytes_to_free are released."""
        entries = sorted(((k, e.memory_usage()) for k, e in self.kv_cache.items()), key=lambda kv: -kv[1])
        freed = 0
        for key, size in entries:
            if freed >= bytes_to_free:
                break
            e = self.kv_cache.pop(key, None)
            if e is not None:
                freed += size
                if hasattr(e, "close"):
                    e.close()
        self.cache_memory_usage = max(0, self.cache_memory_usage - freed)
        return freed

    def clear_kv_cache(self):
        """:1076-1079"""
        for e in self.kv_cache.values():
            if hasattr(e, "close"):
                e.close()
        self.kv_cache.clear()
        self.cache_memory_usage = 0

    def kv_cache_memory_usage(self) -> int:
        """:1082-1084"""
        return self.cache_memory_usage

    def init_kv_cache_dev(self, capacity: int, scheme: int = L.KV_TENSOR_B) -> DeviceKVCacheEntry:
        """lib.rs:958-975 with the entry resident in HBM: empty [layers, 0, hidden], room for `capacity` tokens per layer"""
        q = self.config.quant_config
        pre, dec = (q.prefill_bits, q.decode_bits) if self.config.use_phase_aware_quant else (self.config.kv_quant_bits,) * 2
        return DeviceKVCacheEntry(self._ctx, self.config.num_layers, self.config.hidden_size, capacity, pre, dec, scheme)

    def sample_cached_dev(self, model: QuantizedDiffusionModel, shape, num_steps: Optional[int] = None, cache_id: str = "default",
                          seed: int = 42, x0=None, guard_t0: bool = True, capacity: int = 0, scheme: int = L.KV_TENSOR_B):
        """The cached branch of DiffuseLLM::sample (lib.rs:862-921, :929-936) with x, the cache entry, the keys / values the
        model reads and the noise all on the device: per step the host only does the phase / precision arithmetic
        (:886-903) and enqueues kernels — no byte crosses PCIe between the upload of x0 (or nothing, if x0 is drawn from the
        seeded generator) and the download of the sample.  Noise: the counter-based generator of `sample_seeded`."""
        batch, seq_len = shape
        num_steps = num_steps if num_steps is not None else self.config.num_timesteps
        feat = self.config.hidden_size * seq_len
        ctx, lib = self._ctx, self._ctx._lib
        qc = self.config.quant_config
        n = batch * feat
        with ctx.lock:
            cache = self.kv_cache.get(cache_id)
            if not isinstance(cache, DeviceKVCacheEntry):
                cache = self.init_kv_cache_dev(max(capacity, 1), scheme)
            cache.set_phase(True)                                                              # :866-868
            cap_elems = max(cache.layers * cache.capacity * cache.hidden, 1)
            x_dev, pred_dev = ctx.malloc(n * 4), ctx.malloc(n * 4)
            k_dev, v_dev = ctx.malloc(cap_elems * 4), ctx.malloc(cap_elems * 4)
            try:
                if x0 is not None:
                    ctx.h2d(x_dev, np.ascontiguousarray(x0, np.float32))
                else:
                    ctx._ck(lib.dllm_noise_fill_dev(ctx.h, seed, num_steps, 0, n, x_dev))      # :875-878
                for t in range(num_steps - 1, -1, -1):
                    is_prefill = C.c_int32()
                    target = lib.dllm_progressive_bits(num_steps, t, qc.decode_bits, qc.min_decode_bits, C.byref(is_prefill))
                    cache.set_phase(bool(is_prefill.value))                                    # :886-887
                    if self.config.use_phase_aware_quant and qc.progressive_precision and not is_prefill.value:
                        cache.set_decode_bits(int(target))                                     # :890-903
                    upd = model.update_kv_cache_dev(x_dev, t, batch, feat, cache)              # :907
                    cache.get_dev(k_dev, v_dev)                                                # :913-914
                    model.forward_with_cache_dev(x_dev, t, batch, feat, k_dev, v_dev, cache.len(), pred_dev)   # :910-915
                    if upd[0] == "same":                                                       # :918
                        cache.refresh_dev()
                    elif upd[0] == "update":
                        cache.update_dev(upd[1], upd[2], upd[3])
                    else:
                        cache.append_dev(upd[1], upd[2], upd[3])
                    ctx._ck(lib.dllm_p_sample_seeded_dev(ctx.h, model.h, x_dev, pred_dev, seed, t, batch, feat,
                                                         int(guard_t0), x_dev))                # :921
                self.kv_cache[cache_id] = cache                                                # :929-936
                out = ctx.d2h(x_dev, (batch, feat), np.float32)
            finally:
                ctx.sync()
                for d in (x_dev, pred_dev, k_dev, v_dev):
                    ctx.free(d)
        return out

    def init_kv_cache(self, batch_size: int) -> KVCacheEntry:
        """lib.rs:958-975: empty [layers, 0, hidden] cache with phase-aware bits"""
        shape = (self.config.num_layers, 0, self.config.hidden_size)
        q = self.config.quant_config
        pre, dec = (q.prefill_bits, q.decode_bits) if self.config.use_phase_aware_quant else (self.config.kv_quant_bits,) * 2
        return KVCacheEntry(np.zeros(shape, np.float32), np.zeros(shape, np.float32), pre, dec, self._ctx)

    def sample(self, model: DiffusionModel, shape, num_steps: Optional[int] = None, cache_id: Optional[str] = None,
               x0=None, noises=None, guard_t0: bool = True):
        """lib.rs:853-927.  The reference draws x0 and the per-step noise from an unseeded
        thread_rng (:875-878, :1201); here they are injected (`x0` [batch, hidden*seq], `noises`
        [num_steps, batch, feat], slice t used at timestep t>0) or drawn from numpy when omitted."""
        batch, seq_len = shape
        num_steps = num_steps if num_steps is not None else self.config.num_timesteps
        feat = self.config.hidden_size * seq_len
        rng = np.random.default_rng()
        x = np.ascontiguousarray(x0, np.float32) if x0 is not None else rng.standard_normal((batch, feat)).astype(np.float32)
        if noises is None:
            noises = rng.standard_normal((num_steps, batch, feat)).astype(np.float32)
        noises = np.ascontiguousarray(noises, np.float32)
        use_cache = self.config.use_kv_cache and cache_id is not None

        if not use_cache and isinstance(model, QuantizedDiffusionModel):
            # whole loop on the device: one H2D of x0, per-step noise slices, one D2H of the result
            out = np.empty_like(x)
            with self._ctx.lock:
                self._ctx._ck(self._ctx._lib.dllm_sample(self._ctx.h, model.h, x.ctypes.data, noises.ctypes.data, batch,
                                                         feat, num_steps, int(guard_t0), model.path, out.ctypes.data))
            return out

        cache = None
        if use_cache:
            cache = self.kv_cache.get(cache_id) or self.init_kv_cache(batch)
            cache.set_phase(True)
        qc = self.config.quant_config
        for t in range(num_steps - 1, -1, -1):
            t_array = np.full(batch, t, np.uint64)
            if cache is not None:
                is_prefill = C.c_int32()
                target = self._ctx._lib.dllm_progressive_bits(num_steps, t, qc.decode_bits, qc.min_decode_bits,
                                                              C.byref(is_prefill))              # :886-897
                cache.set_phase(bool(is_prefill.value))
                if (self.config.use_phase_aware_quant and qc.progressive_precision and not is_prefill.value
                        and target != cache.decode_quant_bits):
                    cache.decode_quant_bits = int(target)
                    cache.decode_quantized = None                                             # :900-903
                new_k, new_v = model.update_kv_cache(x, t_array, cache)                        # :907
                pred = model.forward_with_cache(x, t_array, cache.get_keys(), cache.get_values())  # :910-915
                cache.update(new_k, new_v)                                                     # :918
            else:
                pred = model.forward(x, t_array)                                               # :924
            x = self.p_sample(model, x, t_array, pred, noises[t] if t > 0 else None, guard_t0)  # :921/:925
        if cache is not None:
            self.kv_cache[cache_id] = cache                                                    # :929-936
        return x
