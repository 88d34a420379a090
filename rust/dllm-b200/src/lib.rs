//! Safe wrappers over `dllm-b200-sys`: the reference's own seams, backed by libdllm_b200.so.
//!
//! * `quantize_tensor` / `dequantize_tensor`       — diffuse-llm-rs/src/quantization.rs:38,81 (same signatures)
//! * `impl quantization::Quantizer for GpuQuantizer` — quantization/src/quantize.rs:81-90
//! * `impl kvquant::Quantizer for GpuBitQuantizer`   — prefill-kvquant-rs/lib.rs:29-32 (a `Box<dyn Quantizer>` slot)
//! * `impl DiffusionModel for GpuQuantizedModel`     — diffuse-llm-rs/src/lib.rs:748-772
//! * `sample_seeded` / `sample`                      — DiffuseLLM::sample, lib.rs:853-927
//! * `QWeight::{save, load}`                         — the DLLMQW01 packed-weights container
//!
//! NOTE: no Rust toolchain exists in the image this was written in (cargo / rustc / bindgen absent), so this file has not
//! been compiled; the same C ABI is exercised from C++ (diffusion-llm-rs_b200/host/dllm.hpp) and Python (ctypes).
//! Status mapping: 1..=7 -> the same-named `QuantizationError` variant (quantization/src/error.rs:19-40); where the
//! reference panics (`assert!` at quantization.rs:39, Vec index at prefill-kvquant-rs/lib.rs:133) so does the wrapper;
//! 100+ -> `anyhow!` for the `Result<_, anyhow::Error>` methods.
use std::ffi::{CStr, CString};
use std::sync::{Mutex, OnceLock};

use dllm_b200_sys as sys;
use ndarray::{Array1, Array2, Array3, ArrayD, ArrayViewD, IxDyn};

use diffuse_llm_rs::diffuse_llm::{DiffusionModel, KVCacheEntry};
use prefill_kvquant_rs::kvquant;
use quantization::{QuantizationError, QuantizationParams, QuantizationType, QuantizedTensor};

/// One context (CUDA device + stream) behind a mutex: `Send + Sync` for the traits that require it.
pub struct Gpu {
    ctx: Mutex<*mut sys::dllm_ctx>,
}
unsafe impl Send for Gpu {}
unsafe impl Sync for Gpu {}

impl Gpu {
    pub fn new(device: i32) -> anyhow::Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sys::dllm_ctx_create(device, &mut ctx) };
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_ctx_create({device}) failed with status {rc}: no sm_100 GPU — there is no CPU fallback");
        }
        Ok(Gpu { ctx: Mutex::new(ctx) })
    }
    fn with<R>(&self, f: impl FnOnce(*mut sys::dllm_ctx) -> R) -> R {
        let g = self.ctx.lock().unwrap();
        f(*g)
    }
    fn last_error(&self) -> String {
        self.with(|c| unsafe { CStr::from_ptr(sys::dllm_last_error(c)).to_string_lossy().into_owned() })
    }
}
impl Drop for Gpu {
    fn drop(&mut self) {
        unsafe { sys::dllm_ctx_destroy(*self.ctx.lock().unwrap()) }
    }
}

/// Process-wide context on device `LOCAL_RANK` (or 0).
pub fn gpu() -> &'static Gpu {
    static G: OnceLock<Gpu> = OnceLock::new();
    G.get_or_init(|| {
        let dev = std::env::var("LOCAL_RANK").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        Gpu::new(dev).expect("no sm_100 GPU")
    })
}

fn check(g: &Gpu, rc: i32) -> quantization::Result<()> {
    match rc {
        sys::DLLM_OK => Ok(()),
        sys::DLLM_ERR_INVALID_PARAMS => Err(QuantizationError::InvalidParams(g.last_error())),
        sys::DLLM_ERR_UNSUPPORTED => Err(QuantizationError::UnsupportedOperation(g.last_error())),
        sys::DLLM_ERR_SHAPE => Err(QuantizationError::ShapeMismatch(g.last_error())),
        sys::DLLM_ERR_CALIBRATION_REQUIRED => Err(QuantizationError::CalibrationRequired),
        sys::DLLM_ERR_INVALID_DATA_FORMAT | sys::DLLM_ERR_SERIALIZATION => Err(QuantizationError::InvalidDataFormat(g.last_error())),
        sys::DLLM_ERR_INDEX => panic!("{}", g.last_error()), // the reference panics (Vec index out of bounds)
        other => Err(QuantizationError::UnsupportedOperation(format!("dllm status {other}: {}", g.last_error()))),
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// diffuse_llm_rs::quantization — same signatures as diffuse-llm-rs/src/quantization.rs:38,81
// ---------------------------------------------------------------------------------------------------------------------
pub fn quantize_tensor(data: &[f32], bits: u8) -> (Vec<u8>, f32, f32) {
    let g = gpu();
    let mut codes = vec![0u8; data.len()];
    let (mut s, mut z) = (0f32, 0f32);
    let rc = g.with(|c| unsafe { sys::dllm_quantize_tensor(c, data.as_ptr(), data.len(), bits, codes.as_mut_ptr(), &mut s, &mut z) });
    assert!(rc == sys::DLLM_OK, "Bits must be between 1 and 8"); // quantization.rs:39
    (codes, s, z)
}

pub fn dequantize_tensor(data: &[u8], scale: f32, zero_point: f32) -> Vec<f32> {
    let g = gpu();
    let mut out = vec![0f32; data.len()];
    let rc = g.with(|c| unsafe { sys::dllm_dequantize_tensor(c, data.as_ptr(), data.len(), scale, zero_point, out.as_mut_ptr()) });
    assert!(rc == sys::DLLM_OK, "{}", g.last_error());
    out
}

// ---------------------------------------------------------------------------------------------------------------------
// quantization::Quantizer (quantization/src/quantize.rs:81-90)
// ---------------------------------------------------------------------------------------------------------------------
pub struct GpuQuantizer {
    params: QuantizationParams,
}
impl GpuQuantizer {
    /// `DefaultQuantizer::new` hard-codes scale 1.0 / zero-point 0 (quantize.rs:98-108); calibrated parameters
    /// (CalibrationData::compute_params, calibrate.rs:72-110) can be supplied here.
    pub fn new(params: QuantizationParams) -> Self {
        GpuQuantizer { params }
    }
}
impl quantization::Quantizer for GpuQuantizer {
    fn quantize(&self, data: ArrayViewD<f32>, qtype: QuantizationType) -> quantization::Result<QuantizedTensor> {
        let g = gpu();
        let flat: Vec<f32> = data.iter().copied().collect(); // logical order, like quantize.rs:148
        let mut codes = vec![0u8; flat.len()];
        let q = match qtype {
            QuantizationType::Int8 => sys::DLLM_QT_INT8,
            QuantizationType::Int4 => sys::DLLM_QT_INT4,
            QuantizationType::Binary => sys::DLLM_QT_BINARY,
            QuantizationType::Float8 => sys::DLLM_QT_FLOAT8,
        };
        check(g, g.with(|c| unsafe {
            sys::dllm_quantize_a(c, flat.as_ptr(), flat.len(), q, self.params.scale, self.params.zero_point, codes.as_mut_ptr())
        }))?;
        Ok(QuantizedTensor::new(codes, data.shape().to_vec(), self.params.clone()))
    }
    fn dequantize(&self, t: &QuantizedTensor) -> quantization::Result<ArrayD<f32>> {
        let g = gpu();
        let mut out = vec![0f32; t.data.len()];
        check(g, g.with(|c| unsafe {
            sys::dllm_dequantize_a(c, t.data.as_ptr(), t.data.len(), t.params.scale, t.params.zero_point, out.as_mut_ptr())
        }))?;
        ArrayD::from_shape_vec(IxDyn(&t.shape), out).map_err(|e| QuantizationError::ShapeMismatch(e.to_string()))
    }
    fn get_params(&self) -> &QuantizationParams {
        &self.params
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// kvquant::Quantizer (prefill-kvquant-rs/lib.rs:29-32; identical trait at diffusion_prefill/src/prefill_kv.rs:42-45)
// ---------------------------------------------------------------------------------------------------------------------
pub struct GpuBitQuantizer {
    pub scale: f32,
    pub zero_point: f32,
}
impl GpuBitQuantizer {
    /// `BitQuantizer { scale: 1/((1<<bits)-1), zero_point: 0 }`, prefill-kvquant-rs/lib.rs:105-108
    pub fn for_bits(bits: u8) -> Self {
        GpuBitQuantizer { scale: unsafe { sys::dllm_bitquantizer_scale(bits) }, zero_point: 0.0 }
    }
}
impl kvquant::Quantizer for GpuBitQuantizer {
    fn quantize(&self, input: &[f32], bits: u8) -> Vec<u8> {
        let g = gpu();
        let mut codes = vec![0u8; input.len()];
        let rc = g.with(|c| unsafe { sys::dllm_quantize_c(c, input.as_ptr(), input.len(), bits, self.scale, self.zero_point, codes.as_mut_ptr()) });
        assert!(rc == sys::DLLM_OK, "{}", g.last_error());
        codes
    }
    fn dequantize(&self, input: &[u8], _bits: u8) -> Vec<f32> {
        let g = gpu();
        let mut out = vec![0f32; input.len()];
        let rc = g.with(|c| unsafe { sys::dllm_dequantize_cd(c, input.as_ptr(), input.len(), self.scale, self.zero_point, out.as_mut_ptr()) });
        assert!(rc == sys::DLLM_OK, "{}", g.last_error());
        out
    }
}

/// Per-token rows (quantizer D): `KVCache::compress_vector` / `FusionANN::quantize` for a whole batch of rows
/// (diffusion_prefill/src/prefill_kv.rs:104-121, fusion_ann.rs:53-88): row r uses `bits[r % bits.len()]`.
pub fn quantize_rows(rows: &Array2<f32>, bits: &[u8]) -> (Array2<u8>, Vec<f32>, Vec<f32>) {
    let g = gpu();
    let (r, d) = rows.dim();
    let x = rows.as_standard_layout();
    let mut codes = Array2::<u8>::zeros((r, d));
    let (mut sc, mut zp) = (vec![0f32; r], vec![0f32; r]);
    let rc = g.with(|c| unsafe {
        sys::dllm_quantize_d_rows(c, x.as_ptr(), r, d, bits.as_ptr(), bits.len(), codes.as_mut_ptr(), sc.as_mut_ptr(), zp.as_mut_ptr())
    });
    assert!(rc == sys::DLLM_OK, "{}", g.last_error());
    (codes, sc, zp)
}

// ---------------------------------------------------------------------------------------------------------------------
// the quantized linear and the layer stack behind DiffusionModel (diffuse-llm-rs/src/lib.rs:748-813)
// ---------------------------------------------------------------------------------------------------------------------
pub struct QWeight {
    raw: *mut sys::dllm_qweight,
}
unsafe impl Send for QWeight {}
unsafe impl Sync for QWeight {}
impl QWeight {
    /// `weights` is `[input_dim, output_dim]` like SimpleDiffusionModel's (lib.rs:777); `group = 0` is the reference's
    /// per-tensor scale, `group = 128` the grouped extension (quantization/src/types.rs:126).
    pub fn quantize(weights: &Array2<f32>, bits: u8, group: usize, bias: Option<&Array1<f32>>) -> anyhow::Result<Self> {
        let g = gpu();
        let (k, n) = weights.dim();
        let w = weights.as_standard_layout();
        let mut raw = std::ptr::null_mut();
        let b = bias.map(|b| b.as_ptr()).unwrap_or(std::ptr::null());
        let rc = g.with(|c| unsafe { sys::dllm_qweight_quantize(c, w.as_ptr(), k, n, bits, group, b, &mut raw) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_qweight_quantize: status {rc}: {}", g.last_error());
        }
        Ok(QWeight { raw })
    }
    /// y = x · dequant(W) + b  (lib.rs:812 composed with quantization.rs:81-85); `path` = DLLM_PATH_*
    pub fn forward(&self, x: &Array2<f32>, path: i32) -> anyhow::Result<Array2<f32>> {
        let g = gpu();
        let (m, _k) = x.dim();
        let mut n = 0usize;
        unsafe { sys::dllm_qweight_info(self.raw, std::ptr::null_mut(), &mut n, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) };
        let xs = x.as_standard_layout();
        let mut y = Array2::<f32>::zeros((m, n));
        let rc = g.with(|c| unsafe { sys::dllm_qlinear_forward(c, self.raw, xs.as_ptr(), m, y.as_mut_ptr(), path) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_qlinear_forward: status {rc}: {}", g.last_error());
        }
        Ok(y)
    }
    /// Exact integer matmul: y[m,n] = Σ_k xq[m,k]·(q[k,n] − zp); dequantize_tensor ∘ dot is `w_scale * x_scale * y`.
    pub fn matmul_i8(&self, xq: &Array2<i8>) -> anyhow::Result<Array2<i32>> {
        let g = gpu();
        let (m, _k) = xq.dim();
        let mut n = 0usize;
        unsafe { sys::dllm_qweight_info(self.raw, std::ptr::null_mut(), &mut n, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut()) };
        let xs = xq.as_standard_layout();
        let mut y = Array2::<i32>::zeros((m, n));
        let rc = g.with(|c| unsafe { sys::dllm_qlinear_forward_i8(c, self.raw, xs.as_ptr(), m, y.as_mut_ptr()) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_qlinear_forward_i8: status {rc}: {}", g.last_error());
        }
        Ok(y)
    }
    /// DLLMQW01 packed-weights container (what serde on `QuantizedTensor` would carry, bit-packed + CRC)
    pub fn save(&self, path: &str) -> anyhow::Result<()> {
        let g = gpu();
        let p = CString::new(path)?;
        let rc = g.with(|c| unsafe { sys::dllm_qweight_save(c, self.raw, p.as_ptr()) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_qweight_save: status {rc}: {}", g.last_error());
        }
        Ok(())
    }
    pub fn load(path: &str) -> anyhow::Result<Self> {
        let g = gpu();
        let p = CString::new(path)?;
        let mut raw = std::ptr::null_mut();
        let rc = g.with(|c| unsafe { sys::dllm_qweight_load(c, p.as_ptr(), &mut raw) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_qweight_load: status {rc}: {}", g.last_error());
        }
        Ok(QWeight { raw })
    }
}
impl Drop for QWeight {
    fn drop(&mut self) {
        unsafe { sys::dllm_qweight_destroy(self.raw) }
    }
}

/// A stack of quantized linears resident in HBM.  `SimpleDiffusionModel::new(input_dim, output_dim)` (lib.rs:789-801) is the
/// one-layer case: draw `weights` / `bias` as today, quantize once, keep the handle.
pub struct GpuQuantizedModel {
    layers: Vec<QWeight>,
    model: *mut sys::dllm_model,
    hidden: usize,
    /// which kernel family the linears run on (one of the path constants of the -sys crate); `DLLM_PATH_I8` = the int8 denoise mode (per-tensor weights only)
    path: i32,
}
unsafe impl Send for GpuQuantizedModel {}
unsafe impl Sync for GpuQuantizedModel {}
impl GpuQuantizedModel {
    pub fn new(layers: Vec<QWeight>, hidden: usize, num_timesteps: usize, beta_kind: i32, beta_start: f32, beta_end: f32) -> anyhow::Result<Self> {
        let g = gpu();
        let raws: Vec<*mut sys::dllm_qweight> = layers.iter().map(|l| l.raw).collect();
        let mut model = std::ptr::null_mut();
        let rc = g.with(|c| unsafe {
            sys::dllm_model_create(c, hidden, raws.as_ptr(), raws.len(), num_timesteps, beta_kind, beta_start, beta_end, &mut model)
        });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_model_create: status {rc}: {}", g.last_error());
        }
        Ok(GpuQuantizedModel { layers, model, hidden, path: sys::DLLM_PATH_AUTO })
    }
    /// Select the kernel family, e.g. `sys::DLLM_PATH_I8` for the int8 denoise mode (weights quantized per tensor, `group = 0`).
    pub fn with_path(mut self, path: i32) -> Self {
        self.path = path;
        self
    }
    pub fn hidden(&self) -> usize {
        self.hidden
    }
    pub fn n_layers(&self) -> usize {
        self.layers.len()
    }
    /// DiffuseLLM::sample without cache (lib.rs:853-927), noise injected: `noises` is `[num_steps, batch, feat]`.
    pub fn sample(&self, x0: &Array2<f32>, noises: &Array3<f32>, num_steps: usize) -> anyhow::Result<Array2<f32>> {
        let g = gpu();
        let (batch, feat) = x0.dim();
        let mut out = Array2::<f32>::zeros((batch, feat));
        let (x, z) = (x0.as_standard_layout(), noises.as_standard_layout());
        let rc = g.with(|c| unsafe { sys::dllm_sample(c, self.model, x.as_ptr(), z.as_ptr(), batch, feat, num_steps, 1, self.path, out.as_mut_ptr()) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_sample: status {rc}: {}", g.last_error());
        }
        Ok(out)
    }
    /// The same loop with the noise drawn on the device from the counter-based generator (nothing uploaded per step; the
    /// step is replayed from one CUDA graph).  42 is the one seed the reference uses (examples/diffusion_example.rs:69).
    pub fn sample_seeded(&self, batch: usize, feat: usize, num_steps: usize, seed: u64) -> anyhow::Result<Array2<f32>> {
        let g = gpu();
        let mut out = Array2::<f32>::zeros((batch, feat));
        let rc = g.with(|c| unsafe {
            sys::dllm_sample_seeded(c, self.model, std::ptr::null(), seed, batch, feat, num_steps, 1, self.path, 1, out.as_mut_ptr())
        });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_sample_seeded: status {rc}: {}", g.last_error());
        }
        Ok(out)
    }
}
impl Drop for GpuQuantizedModel {
    fn drop(&mut self) {
        unsafe { sys::dllm_model_destroy(self.model) }
    }
}

#[async_trait::async_trait]
impl DiffusionModel for GpuQuantizedModel {
    fn forward(&self, x: &Array2<f32>, t: &Array1<usize>) -> Array2<f32> {
        let g = gpu();
        let (batch, feat) = x.dim();
        let xs = x.as_standard_layout();
        let mut out = Array2::<f32>::zeros((batch, feat));
        let rc = g.with(|c| unsafe {
            sys::dllm_model_forward(c, self.model, xs.as_ptr(), t.as_ptr(), batch, feat, out.as_mut_ptr(), self.path)
        });
        assert_eq!(rc, sys::DLLM_OK, "{}", g.last_error()); // forward() is infallible in the trait
        out
    }
    fn forward_with_cache(&self, x: &Array2<f32>, t: &Array1<usize>, _keys: &Array3<f32>, _values: &Array3<f32>) -> Result<Array2<f32>, anyhow::Error> {
        Ok(self.forward(x, t)) // lib.rs:815-824: the cache is ignored
    }
    fn update_kv_cache(&self, _x: &Array2<f32>, _t: &Array1<usize>, cache: &KVCacheEntry) -> Result<(Array3<f32>, Array3<f32>), anyhow::Error> {
        Ok((cache.keys.clone(), cache.values.clone())) // lib.rs:826-835
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// KV cache entry resident in HBM (QuantizedKVCacheEntry, quantization.rs:129-176; growth: lib.rs:246-276, :913-918)
// ---------------------------------------------------------------------------------------------------------------------
pub struct GpuKvEntry {
    raw: *mut sys::dllm_kv,
    shape: (usize, usize, usize),
}
unsafe impl Send for GpuKvEntry {}
unsafe impl Sync for GpuKvEntry {}
impl GpuKvEntry {
    /// `QuantizedKVCacheEntry::new(keys, values, bits)`: each tensor quantized as one (scheme = DLLM_KV_TENSOR_B)
    pub fn new(keys: &Array3<f32>, values: &Array3<f32>, bits: u8, scheme: i32) -> anyhow::Result<Self> {
        let g = gpu();
        let (l, s, h) = keys.dim();
        let (k, v) = (keys.as_standard_layout(), values.as_standard_layout());
        let mut raw = std::ptr::null_mut();
        let rc = g.with(|c| unsafe { sys::dllm_kv_quantize(c, k.as_ptr(), v.as_ptr(), l, s, h, bits, scheme, &mut raw) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_kv_quantize: status {rc}: {}", g.last_error());
        }
        Ok(GpuKvEntry { raw, shape: (l, s, h) })
    }
    /// An empty entry that grows: only the new `[layers, t_new, hidden]` slice crosses the boundary per step.
    pub fn with_capacity(layers: usize, capacity: usize, hidden: usize, bits: u8, scheme: i32) -> anyhow::Result<Self> {
        let g = gpu();
        let mut raw = std::ptr::null_mut();
        let rc = g.with(|c| unsafe { sys::dllm_kv_create(c, layers, capacity, hidden, bits, scheme, &mut raw) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_kv_create: status {rc}: {}", g.last_error());
        }
        Ok(GpuKvEntry { raw, shape: (layers, 0, hidden) })
    }
    pub fn append(&mut self, new_keys: &Array3<f32>, new_values: &Array3<f32>) -> anyhow::Result<()> {
        let g = gpu();
        let t_new = new_keys.dim().1;
        let (k, v) = (new_keys.as_standard_layout(), new_values.as_standard_layout());
        let rc = g.with(|c| unsafe { sys::dllm_kv_append(c, self.raw, k.as_ptr(), v.as_ptr(), t_new) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_kv_append: status {rc}: {}", g.last_error());
        }
        self.shape.1 += t_new;
        Ok(())
    }
    pub fn dequantize(&self) -> anyhow::Result<(Array3<f32>, Array3<f32>)> {
        let g = gpu();
        let mut k = Array3::<f32>::zeros(self.shape);
        let mut v = Array3::<f32>::zeros(self.shape);
        let rc = g.with(|c| unsafe { sys::dllm_kv_dequantize(c, self.raw, k.as_mut_ptr(), v.as_mut_ptr()) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_kv_dequantize: status {rc}: {}", g.last_error());
        }
        Ok((k, v))
    }
    /// KVCacheEntry::memory_usage accounting, lib.rs:279-302
    pub fn memory_usage(&self) -> usize {
        unsafe { sys::dllm_kv_memory_usage(self.raw) }
    }
}
impl Drop for GpuKvEntry {
    fn drop(&mut self) {
        unsafe { sys::dllm_kv_destroy(self.raw) }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Phase-aware cache entry resident in HBM (KVCacheEntry, lib.rs:122-313): f32 keys / values + a prefill- and a decode-precision
// quantized copy, all on the device.  The device-pointer methods are what a GPU-resident sampling loop uses (lib.rs:885-921
// without PCIe traffic); the ndarray methods stage through host memory for drop-in use.
// ---------------------------------------------------------------------------------------------------------------------
pub struct GpuKvCacheEntry {
    raw: *mut sys::dllm_kvcache,
    layers: usize,
    hidden: usize,
}
unsafe impl Send for GpuKvCacheEntry {}
unsafe impl Sync for GpuKvCacheEntry {}
impl GpuKvCacheEntry {
    /// `KVCacheEntry::new` with room for `capacity` tokens per layer (scheme = DLLM_KV_TENSOR_B is the reference's quantizer)
    pub fn with_capacity(layers: usize, hidden: usize, capacity: usize, prefill_bits: u8, decode_bits: u8, scheme: i32) -> anyhow::Result<Self> {
        let g = gpu();
        let mut raw = std::ptr::null_mut();
        let rc = g.with(|c| unsafe { sys::dllm_kvcache_create(c, layers, hidden, capacity, prefill_bits, decode_bits, scheme, &mut raw) });
        if rc != sys::DLLM_OK {
            anyhow::bail!("dllm_kvcache_create: status {rc}: {}", g.last_error());
        }
        Ok(GpuKvCacheEntry { raw, layers, hidden })
    }
    fn info(&self) -> (usize, bool, u8, usize) {
        let (mut s, mut p, mut b, mut m) = (0usize, 0i32, 0u8, 0usize);
        unsafe { sys::dllm_kvcache_info(self.raw, &mut s, &mut p, &mut b, &mut m) };
        (s, p != 0, b, m)
    }
    pub fn len(&self) -> usize { self.info().0 }
    pub fn is_empty(&self) -> bool { self.len() == 0 }
    pub fn get_current_quant_bits(&self) -> u8 { self.info().2 }
    pub fn memory_usage(&self) -> usize { self.info().3 }
    /// `transition_phase` / `set_phase` (lib.rs:207-238)
    pub fn set_phase(&mut self, is_prefill: bool) -> anyhow::Result<()> {
        let g = gpu();
        let rc = g.with(|c| unsafe { sys::dllm_kvcache_set_phase(c, self.raw, is_prefill as i32) });
        if rc != sys::DLLM_OK { anyhow::bail!("dllm_kvcache_set_phase: status {rc}: {}", g.last_error()); }
        Ok(())
    }
    /// `decode_quant_bits = bits; decode_quantized = None` (lib.rs:899-903)
    pub fn set_decode_bits(&mut self, bits: u8) -> anyhow::Result<()> {
        let g = gpu();
        let rc = g.with(|c| unsafe { sys::dllm_kvcache_set_decode_bits(c, self.raw, bits) });
        if rc != sys::DLLM_OK { anyhow::bail!("dllm_kvcache_set_decode_bits: status {rc}: {}", g.last_error()); }
        Ok(())
    }
    /// `update` on device tensors `[layers, seq, hidden]` (lib.rs:246-276); null pointers + the current length re-quantize the
    /// entry's own tensors
    pub unsafe fn update_dev(&mut self, keys_dev: *const f32, values_dev: *const f32, seq: usize) -> anyhow::Result<()> {
        let g = gpu();
        let rc = g.with(|c| sys::dllm_kvcache_update_dev(c, self.raw, keys_dev, values_dev, seq));
        if rc != sys::DLLM_OK { anyhow::bail!("dllm_kvcache_update_dev: status {rc}: {}", g.last_error()); }
        Ok(())
    }
    /// only `t_new` tokens per layer are new (per-token / fixed-scale entries quantize just those)
    pub unsafe fn append_dev(&mut self, keys_new_dev: *const f32, values_new_dev: *const f32, t_new: usize) -> anyhow::Result<()> {
        let g = gpu();
        let rc = g.with(|c| sys::dllm_kvcache_append_dev(c, self.raw, keys_new_dev, values_new_dev, t_new));
        if rc != sys::DLLM_OK { anyhow::bail!("dllm_kvcache_append_dev: status {rc}: {}", g.last_error()); }
        Ok(())
    }
    /// `get_keys` / `get_values` decoded straight into the consumer's device buffers (either may be null)
    pub unsafe fn get_dev(&self, keys_out_dev: *mut f32, values_out_dev: *mut f32) -> anyhow::Result<()> {
        let g = gpu();
        let rc = g.with(|c| sys::dllm_kvcache_get_dev(c, self.raw, keys_out_dev, values_out_dev));
        if rc != sys::DLLM_OK { anyhow::bail!("dllm_kvcache_get_dev: status {rc}: {}", g.last_error()); }
        Ok(())
    }
    /// drop-in `update(new_keys, new_values)` from host arrays
    pub fn update(&mut self, new_keys: &Array3<f32>, new_values: &Array3<f32>) -> anyhow::Result<()> {
        let g = gpu();
        let (l, s, h) = new_keys.dim();
        anyhow::ensure!(l == self.layers && h == self.hidden, "shape mismatch");
        let (k, v) = (new_keys.as_standard_layout(), new_values.as_standard_layout());
        let bytes = l * s * h * 4;
        g.with(|c| unsafe {
            let (mut dk, mut dv) = (std::ptr::null_mut(), std::ptr::null_mut());
            let mut rc = sys::dllm_malloc(c, bytes.max(4), &mut dk);
            if rc == sys::DLLM_OK { rc = sys::dllm_malloc(c, bytes.max(4), &mut dv); }
            if rc == sys::DLLM_OK { rc = sys::dllm_memcpy_h2d(c, dk, k.as_ptr() as *const _, bytes); }
            if rc == sys::DLLM_OK { rc = sys::dllm_memcpy_h2d(c, dv, v.as_ptr() as *const _, bytes); }
            if rc == sys::DLLM_OK { rc = sys::dllm_kvcache_update_dev(c, self.raw, dk as *const f32, dv as *const f32, s); }
            if rc == sys::DLLM_OK { rc = sys::dllm_ctx_sync(c); }
            sys::dllm_free(c, dk);
            sys::dllm_free(c, dv);
            if rc != sys::DLLM_OK { anyhow::bail!("GpuKvCacheEntry::update: status {rc}"); }
            Ok(())
        })
    }
}
impl Drop for GpuKvCacheEntry {
    fn drop(&mut self) {
        unsafe { sys::dllm_kvcache_destroy(self.raw) }
    }
}
