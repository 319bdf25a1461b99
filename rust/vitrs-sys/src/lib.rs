//! Thin FFI crate over `libvitrs.so` (include/vitrs.h) — the boundary the north-star asks for.
//!
//! `ffi` holds one `extern "C"` declaration per exported symbol, generated from the header
//! (scripts/gen_rust_ffi.py; tests/test_abi.py keeps it current).  This file adds the mirror of
//! the reference's model surface on top: `ViT::{build_from_checkpoint, forward, backward}`
//! (rusty_vit.rs:79,269,354), `optimizer_step` (rusty_vit.rs:949) and the train loop's step
//! (train_vit.rs:65-86), with the same names and argument meaning, device-resident.
//!
//! Source only: this image has no Rust toolchain, so the same ABI is exercised from C++
//! (csrc/model.cu) and from ctypes (vit.rs_b200/__init__.py, tests/).  Pointers passed to
//! `forward` / `train_step` are DEVICE pointers; the `_host` calls take host slices.
pub mod ffi;
pub use ffi::*;
use std::ffi::{CStr, CString};
use std::os::raw::c_int;

pub const VITRS_OK: c_int = 0;
pub const VITRS_ERR_CUDA: c_int = -1;
pub const VITRS_ERR_ARG: c_int = -2;
pub const VITRS_ERR_UNSUPPORTED: c_int = -3;
pub const VITRS_ERR_NCCL: c_int = -4;
pub const VITRS_MODE_F32: c_int = 0;
pub const VITRS_MODE_BF16: c_int = 1;
pub const VITRS_NUM_PARAMETER_TENSORS: usize = 20;
pub const VITRS_NUM_ACTIVATION_TENSORS: usize = 23;

/// AdamW hyper-parameters of `ViT::update` (DEVIATIONS D8; the reference's `optimizer_step` is plain SGD).
#[derive(Clone, Copy, Debug)]
pub struct AdamW { pub lr: f32, pub beta1: f32, pub beta2: f32, pub eps: f32, pub weight_decay: f32 }
impl Default for AdamW {
    fn default() -> Self { AdamW { lr: 1e-3, beta1: 0.9, beta2: 0.999, eps: 1e-8, weight_decay: 0.0 } }
}

/// Mirror of the reference's `ViT` (rusty_vit.rs:63-76): same method names, device-resident.
pub struct ViT {
    ctx: *mut vitrs_ctx,
    model: *mut vitrs_model,
    pub config: vitrs_config,
    pub mean_loss: f32,
}

unsafe fn err(ctx: *mut vitrs_ctx) -> String { CStr::from_ptr(vitrs_last_error(ctx)).to_string_lossy().into_owned() }

macro_rules! check {
    ($self:ident, $call:expr) => {
        if unsafe { $call } != VITRS_OK { return Err(unsafe { err($self.ctx) }); }
    };
}

impl ViT {
    /// Allocate a model for `config` on `device` and initialise it (rusty_vit.rs:864 `init_parameters`).
    pub fn new(config: vitrs_config, device: i32, max_batch: i32, mode: c_int, seed: u64) -> Result<ViT, String> {
        unsafe {
            let mut ctx = std::ptr::null_mut();
            if vitrs_ctx_create(&mut ctx, device) != VITRS_OK { return Err("no sm_100 device (there is no CPU fallback)".into()); }
            let mut model = std::ptr::null_mut();
            if vitrs_model_create(ctx, &config, max_batch, mode, &mut model) != VITRS_OK { let e = err(ctx); vitrs_ctx_destroy(ctx); return Err(e); }
            if vitrs_model_init_parameters(model, seed, 0) != VITRS_OK { return Err(err(ctx)); }
            Ok(ViT { ctx, model, config, mean_loss: -1.0 })
        }
    }
    /// `ViT::build_from_checkpoint` (rusty_vit.rs:79): allocate for `config`, then read the llm.c-style file.
    pub fn build_from_checkpoint(config: vitrs_config, max_batch: i32, path: &str) -> Result<ViT, String> {
        let v = ViT::new(config, 0, max_batch, VITRS_MODE_BF16, 0)?;
        let c = CString::new(path).map_err(|e| e.to_string())?;
        check!(v, vitrs_model_load_checkpoint(v.model, c.as_ptr()));
        Ok(v)
    }
    pub fn save_checkpoint(&mut self, path: &str) -> Result<(), String> {
        let c = CString::new(path).map_err(|e| e.to_string())?;
        check!(self, vitrs_model_save_checkpoint(self.model, c.as_ptr()));
        Ok(())
    }
    /// `vit.forward(inputs, targets, b, t)` (rusty_vit.rs:269): device images [b,3,H,W] fp32, device labels [b] (null => logits only).
    pub fn forward(&mut self, images: *const f32, targets: *const c_int, b: i32) -> Result<(), String> {
        check!(self, vitrs_model_forward(self.model, images, targets, b));
        check!(self, vitrs_model_mean_loss(self.model, &mut self.mean_loss));
        Ok(())
    }
    /// `vit.backward()` (rusty_vit.rs:354); gradients accumulate, so zero them once per step first.
    pub fn backward(&mut self) -> Result<(), String> { check!(self, vitrs_model_backward(self.model)); Ok(()) }
    pub fn zero_grad(&mut self) -> Result<(), String> { check!(self, vitrs_model_zero_grad(self.model)); Ok(()) }
    /// AdamW form of `optimizer_step` (under data parallel the gradient exchange has already been queued by `backward`).
    pub fn update(&mut self, o: &AdamW) -> Result<(), String> {
        check!(self, vitrs_model_update(self.model, o.lr, o.beta1, o.beta2, o.eps, o.weight_decay));
        Ok(())
    }
    /// One step of the train loop (train_vit.rs:65-86: forward, zero_grad, backward, update) from HOST slices;
    /// returns the mean loss.  `images` is [b,3,H,W] fp32, `labels` is [b].
    pub fn train_step_host(&mut self, images: &[f32], labels: &[c_int], o: &AdamW) -> Result<f32, String> {
        let b = labels.len() as c_int;
        let per = 3 * (self.config.image_size as usize).pow(2);
        if images.len() != per * labels.len() { return Err("images.len() != b * 3 * H * W".into()); }
        check!(self, vitrs_model_train_step_host(self.model, images.as_ptr(), labels.as_ptr(), b, o.lr, o.beta1, o.beta2, o.eps,
                                                   o.weight_decay, &mut self.mean_loss));
        Ok(self.mean_loss)
    }
    /// The same from raw dataset bytes (uint8; layout 0 = NCHW, 1 = NHWC), normalised on the device.
    pub fn train_step_host_u8(&mut self, images: &[u8], layout: c_int, labels: &[c_int], o: &AdamW) -> Result<f32, String> {
        check!(self, vitrs_model_train_step_host_u8(self.model, images.as_ptr(), layout, labels.as_ptr(), labels.len() as c_int, o.lr, o.beta1,
                                                      o.beta2, o.eps, o.weight_decay, &mut self.mean_loss));
        Ok(self.mean_loss)
    }
    /// Join a data-parallel job: `id128` is the 128-byte NCCL unique id made by rank 0 (`unique_id`) and broadcast by the host.
    pub fn init_data_parallel(&mut self, id128: &[u8; 128], rank: i32, world: i32, global_batch: i32) -> Result<(), String> {
        check!(self, vitrs_comm_init(self.ctx, id128.as_ptr() as *const _, rank, world));
        check!(self, vitrs_model_set_dloss_scale(self.model, 1.0 / global_batch as f32));
        Ok(())
    }
    pub fn unique_id(&mut self) -> Result<[u8; 128], String> {
        let mut id = [0u8; 128];
        check!(self, vitrs_comm_unique_id(self.ctx, id.as_mut_ptr() as *mut _));
        Ok(id)
    }
    /// `ncclCommGetAsyncError`: Err when a peer or the fabric has failed since the last call.
    pub fn comm_health(&mut self) -> Result<(), String> {
        let mut r: c_int = 0;
        check!(self, vitrs_comm_async_error(self.ctx, &mut r));
        Ok(())
    }
    /// ZeRO-1: shard the fp32 master weights and AdamW moments of the GEMM weights over the ranks.
    pub fn enable_zero1(&mut self) -> Result<(), String> { check!(self, vitrs_model_enable_zero1(self.model)); Ok(()) }
    pub fn num_parameters(&self) -> usize { unsafe { vitrs_model_num_parameters(self.model) } }
    pub fn synchronize(&mut self) -> Result<(), String> { check!(self, vitrs_ctx_synchronize(self.ctx)); Ok(()) }
    pub fn raw(&self) -> (*mut vitrs_ctx, *mut vitrs_model) { (self.ctx, self.model) }
}

/// `optimizer_step(model, lr)` (rusty_vit.rs:949): SGD over the flat parameter buffer.
pub fn optimizer_step(model: &mut ViT, lr: f32) -> Result<(), String> {
    check!(model, vitrs_model_optimizer_step(model.model, lr));
    Ok(())
}

/// Device bytes a model of `config` at `max_batch` images will allocate, by what they hold (host arithmetic, no device):
/// size a deployment for the 180 GB of a B200 before creating anything.  `world` / `zero1` describe the data-parallel setup.
pub fn model_footprint(config: &vitrs_config, max_batch: i32, mode: c_int, world: i32, zero1: bool) -> Result<vitrs_footprint, String> {
    let mut f = vitrs_footprint::default();
    let rc = unsafe { vitrs_model_footprint(config, max_batch, mode, world, zero1 as c_int, &mut f) };
    if rc != 0 { return Err(format!("vitrs_model_footprint: invalid configuration ({})", rc)); }
    Ok(f)
}

/// What the library launches for a dense bf16 GEMM of these extents (kernel family, tile, CTA pair, split-K, grid).
pub fn gemm_plan(m: i32, n: i32, k: i32, a_mn_major: bool, b_mn_major: bool, epilogue: c_int, sm_count: i32, flags: c_int) -> Result<vitrs_gemm_plan_t, String> {
    let mut p = vitrs_gemm_plan_t::default();
    let rc = unsafe { vitrs_gemm_plan(m, n, k, a_mn_major as c_int, b_mn_major as c_int, epilogue, sm_count, flags, &mut p) };
    if rc != 0 { return Err(format!("vitrs_gemm_plan: invalid extents ({})", rc)); }
    Ok(p)
}

impl Drop for ViT {
    fn drop(&mut self) { unsafe { vitrs_model_destroy(self.model); vitrs_ctx_destroy(self.ctx); } }
}
