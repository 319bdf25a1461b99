//! `extern "C"` declarations for `libvitrs.so` (include/vitrs.h) — the thin FFI crate the
//! north-star asks for.  Source only: this image has no Rust toolchain, so the same ABI is
//! exercised from C++ (csrc/model.cu) and from ctypes (vit.rs_b200/__init__.py, tests/).
//!
//! Every function mirrors one reference item; the reference line is given beside it
//! (tv = train_vit.rs, rv = rusty_vit.rs).  Pointers are DEVICE pointers, calls are
//! asynchronous on the context's stream, the return value is 0 or a negative status.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct vitrs_ctx { _private: [u8; 0] }
#[repr(C)]
pub struct vitrs_model { _private: [u8; 0] }
pub type vitrs_bf16 = u16;

pub const VITRS_OK: c_int = 0;
pub const VITRS_MODE_F32: c_int = 0;
pub const VITRS_MODE_BF16: c_int = 1;

/// `ViTConfig` (rv:10-16 / tv:56-63) plus the ViT fields (DEVIATIONS D7).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct vitrs_config {
    pub max_seq_len: c_int,
    pub vocab_size: c_int,
    pub num_layers: c_int,
    pub num_heads: c_int,
    pub channels: c_int,
    pub image_size: c_int,
    pub patch_size: c_int,
    pub num_classes: c_int,
    pub causal: c_int,
}

extern "C" {
    pub fn vitrs_ctx_create(out: *mut *mut vitrs_ctx, device: c_int) -> c_int;
    pub fn vitrs_ctx_destroy(ctx: *mut vitrs_ctx) -> c_int;
    pub fn vitrs_ctx_synchronize(ctx: *mut vitrs_ctx) -> c_int;
    pub fn vitrs_last_error(ctx: *mut vitrs_ctx) -> *const c_char;
    pub fn vitrs_malloc(ctx: *mut vitrs_ctx, ptr: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn vitrs_free(ctx: *mut vitrs_ctx, ptr: *mut c_void) -> c_int;
    pub fn vitrs_memcpy_h2d(ctx: *mut vitrs_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn vitrs_memcpy_d2h(ctx: *mut vitrs_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;

    // ---- L1 operators, fp32 verify mode (reference line each replaces) ----
    pub fn vitrs_residual_forward_f32(ctx: *mut vitrs_ctx, out: *mut f32, inp1: *const f32, inp2: *const f32, n: c_int) -> c_int; // tv:376
    pub fn vitrs_matmul_forward_f32(ctx: *mut vitrs_ctx, out: *mut f32, inp: *const f32, weight: *const f32, bias: *const f32,
                                    b: c_int, t: c_int, c: c_int, oc: c_int) -> c_int; // tv:384
    pub fn vitrs_attention_forward_f32(ctx: *mut vitrs_ctx, out: *mut f32, preatt: *mut f32, att: *mut f32, inp: *const f32,
                                       b: c_int, t: c_int, c: c_int, nh: c_int, causal: c_int) -> c_int; // tv:400
    pub fn vitrs_layernorm_forward_f32(ctx: *mut vitrs_ctx, out: *mut f32, mean: *mut f32, rstd: *mut f32, inp: *const f32,
                                       weight: *const f32, bias: *const f32, b: c_int, t: c_int, c: c_int) -> c_int; // tv:453
    pub fn vitrs_gelu_forward_f32(ctx: *mut vitrs_ctx, out: *mut f32, inp: *const f32, n: c_int) -> c_int; // tv:482
    pub fn vitrs_softmax_forward_f32(ctx: *mut vitrs_ctx, probs: *mut f32, logits: *const f32, b: c_int, t: c_int, v: c_int) -> c_int; // tv:493
    pub fn vitrs_residual_backward_f32(ctx: *mut vitrs_ctx, dinp1: *mut f32, dinp2: *mut f32, dout: *const f32, n: c_int) -> c_int; // tv:521
    pub fn vitrs_matmul_backward_f32(ctx: *mut vitrs_ctx, dinp: *mut f32, dweight: *mut f32, dbias: *mut f32, dout: *const f32,
                                     inp: *const f32, weight: *const f32, b: c_int, t: c_int, c: c_int, oc: c_int) -> c_int; // tv:530
    pub fn vitrs_attention_backward_f32(ctx: *mut vitrs_ctx, dinp: *mut f32, dpreatt: *mut f32, datt: *mut f32, dout: *const f32,
                                        inp: *const f32, att: *const f32, b: c_int, t: c_int, c: c_int, nh: c_int, causal: c_int) -> c_int; // tv:559
    pub fn vitrs_layernorm_backward_f32(ctx: *mut vitrs_ctx, dinp: *mut f32, dweight: *mut f32, dbias: *mut f32, dout: *const f32,
                                        inp: *const f32, weight: *const f32, mean: *const f32, rstd: *const f32,
                                        b: c_int, t: c_int, c: c_int) -> c_int; // tv:603
    pub fn vitrs_gelu_backward_f32(ctx: *mut vitrs_ctx, dinp: *mut f32, inp: *const f32, dout: *const f32, n: c_int) -> c_int; // tv:639

    // ---- bf16 production mode: same names with _bf16 (activations / weights bf16, stats and parameter gradients fp32) ----
    pub fn vitrs_matmul_forward_bf16(ctx: *mut vitrs_ctx, out: *mut vitrs_bf16, inp: *const vitrs_bf16, weight: *const vitrs_bf16,
                                     bias: *const f32, b: c_int, t: c_int, c: c_int, oc: c_int) -> c_int;
    pub fn vitrs_matmul_backward_bf16(ctx: *mut vitrs_ctx, dinp: *mut vitrs_bf16, dweight: *mut f32, dbias: *mut f32,
                                      dout: *const vitrs_bf16, inp: *const vitrs_bf16, weight: *const vitrs_bf16,
                                      b: c_int, t: c_int, c: c_int, oc: c_int) -> c_int;
    /// one GEMM with a fused epilogue: 1 bias, 2 bias + GELU (second output d2), 3 bias + residual (aux), 4 gelu'(aux)
    pub fn vitrs_gemm_bf16_fused(ctx: *mut vitrs_ctx, d: *mut vitrs_bf16, d2: *mut vitrs_bf16, aux: *const vitrs_bf16,
                                 bias: *const f32, a_colsum: *mut f32, a: *const vitrs_bf16, b: *const vitrs_bf16,
                                 m: c_int, n: c_int, k: c_int, lda: c_int, ldb: c_int, ldd: c_int,
                                 a_mn_major: c_int, b_mn_major: c_int, epilogue: c_int) -> c_int;
    pub fn vitrs_attention_forward_bf16(ctx: *mut vitrs_ctx, out: *mut vitrs_bf16, lse: *mut f32, inp: *const vitrs_bf16,
                                        b: c_int, t: c_int, c: c_int, nh: c_int, causal: c_int) -> c_int;
    pub fn vitrs_attention_backward_bf16(ctx: *mut vitrs_ctx, dinp: *mut vitrs_bf16, dout: *const vitrs_bf16, out: *const vitrs_bf16,
                                         lse: *const f32, inp: *const vitrs_bf16, b: c_int, t: c_int, c: c_int, nh: c_int,
                                         causal: c_int) -> c_int;
    pub fn vitrs_layernorm_forward_bf16(ctx: *mut vitrs_ctx, out: *mut vitrs_bf16, mean: *mut f32, rstd: *mut f32,
                                        inp: *const vitrs_bf16, weight: *const f32, bias: *const f32, b: c_int, t: c_int, c: c_int) -> c_int;
    pub fn vitrs_layernorm_backward_bf16(ctx: *mut vitrs_ctx, dinp: *mut vitrs_bf16, dweight: *mut f32, dbias: *mut f32,
                                         dout: *const vitrs_bf16, inp: *const vitrs_bf16, weight: *const f32, mean: *const f32,
                                         rstd: *const f32, b: c_int, t: c_int, c: c_int) -> c_int;

    // ---- optimiser: optimizer_step (rv:949) and its AdamW form ----
    pub fn vitrs_sgd_step(ctx: *mut vitrs_ctx, params: *mut f32, grads: *const f32, n: usize, lr: f32, shadow: *mut vitrs_bf16) -> c_int;
    pub fn vitrs_adamw_step(ctx: *mut vitrs_ctx, params: *mut f32, grads: *const f32, m: *mut f32, v: *mut f32, n: usize, lr: f32,
                            beta1: f32, beta2: f32, eps: f32, weight_decay: f32, step: c_int, shadow: *mut vitrs_bf16) -> c_int;

    // ---- L2 model: struct ViT / impl ViT (rv:63-450) ----
    pub fn vitrs_model_create(ctx: *mut vitrs_ctx, cfg: *const vitrs_config, max_batch: c_int, mode: c_int, out: *mut *mut vitrs_model) -> c_int;
    pub fn vitrs_model_destroy(m: *mut vitrs_model) -> c_int;
    pub fn vitrs_model_init_parameters(m: *mut vitrs_model, seed: u64, init_mode: c_int) -> c_int; // rv:864
    pub fn vitrs_model_load_checkpoint(m: *mut vitrs_model, path: *const c_char) -> c_int; // rv:79
    pub fn vitrs_model_save_checkpoint(m: *mut vitrs_model, path: *const c_char) -> c_int; // rv:912
    pub fn vitrs_model_num_parameters(m: *mut vitrs_model) -> usize;
    pub fn vitrs_model_forward(m: *mut vitrs_model, images: *const f32, labels: *const c_int, b: c_int) -> c_int; // rv:269
    pub fn vitrs_model_zero_grad(m: *mut vitrs_model) -> c_int;
    pub fn vitrs_model_backward(m: *mut vitrs_model) -> c_int; // rv:354
    pub fn vitrs_model_optimizer_step(m: *mut vitrs_model, lr: f32) -> c_int; // rv:949
    pub fn vitrs_model_update(m: *mut vitrs_model, lr: f32, beta1: f32, beta2: f32, eps: f32, weight_decay: f32) -> c_int;
    pub fn vitrs_model_mean_loss(m: *mut vitrs_model, out: *mut f32) -> c_int; // rv:75
    pub fn vitrs_model_train_step_host(m: *mut vitrs_model, h_images: *const f32, h_labels: *const c_int, b: c_int, lr: f32, beta1: f32,
                                       beta2: f32, eps: f32, weight_decay: f32, loss_out: *mut f32) -> c_int;
    // raw dataset images (uint8; layout 0 = NCHW, 1 = NHWC), normalised on the device inside the patch embedding
    pub fn vitrs_model_set_input_norm(m: *mut vitrs_model, mean3: *const f32, std3: *const f32) -> c_int;
    pub fn vitrs_model_forward_u8(m: *mut vitrs_model, images: *const u8, layout: c_int, labels: *const c_int, b: c_int) -> c_int;
    pub fn vitrs_model_train_step_u8(m: *mut vitrs_model, images: *const u8, layout: c_int, labels: *const c_int, b: c_int, lr: f32,
                                     beta1: f32, beta2: f32, eps: f32, weight_decay: f32) -> c_int;
    pub fn vitrs_model_prefetch_host_u8(m: *mut vitrs_model, h_images: *const u8, h_labels: *const c_int, b: c_int) -> c_int;
    pub fn vitrs_model_train_step_host_u8(m: *mut vitrs_model, h_images: *const u8, layout: c_int, h_labels: *const c_int, b: c_int,
                                          lr: f32, beta1: f32, beta2: f32, eps: f32, weight_decay: f32, loss_out: *mut f32) -> c_int;
    pub fn vitrs_model_param_view(m: *mut vitrs_model, which: c_int, tensor: c_int, ptr: *mut *mut f32, count: *mut usize) -> c_int;
}

/// Safe-ish mirror of the reference's `ViT` (rv:63-76): same method names, device-resident.
pub struct ViT {
    ctx: *mut vitrs_ctx,
    model: *mut vitrs_model,
    pub config: vitrs_config,
    pub mean_loss: f32,
}

impl ViT {
    /// `ViT::build_from_checkpoint` (rv:79): allocate for `config`, then read the llm.c-style file.
    pub fn build_from_checkpoint(config: vitrs_config, max_batch: i32, path: &str) -> Result<ViT, String> {
        unsafe {
            let mut ctx = std::ptr::null_mut();
            if vitrs_ctx_create(&mut ctx, 0) != VITRS_OK { return Err("no sm_100 device (there is no CPU fallback)".into()); }
            let mut model = std::ptr::null_mut();
            if vitrs_model_create(ctx, &config, max_batch, VITRS_MODE_BF16, &mut model) != VITRS_OK { return Err(err(ctx)); }
            let c = std::ffi::CString::new(path).unwrap();
            if vitrs_model_load_checkpoint(model, c.as_ptr()) != VITRS_OK { return Err(err(ctx)); }
            Ok(ViT { ctx, model, config, mean_loss: -1.0 })
        }
    }
    /// `vit.forward(inputs, targets, b, t)` (rv:269): device images [b,3,H,W] fp32, device labels [b] (null => logits only).
    pub fn forward(&mut self, images: *const f32, targets: *const c_int, b: i32) -> Result<(), String> {
        unsafe {
            if vitrs_model_forward(self.model, images, targets, b) != VITRS_OK { return Err(err(self.ctx)); }
            if vitrs_model_mean_loss(self.model, &mut self.mean_loss) != VITRS_OK { return Err(err(self.ctx)); }
        }
        Ok(())
    }
    /// `vit.backward()` (rv:354); gradients accumulate, so zero them once per step first.
    pub fn backward(&mut self) -> Result<(), String> {
        unsafe { if vitrs_model_backward(self.model) != VITRS_OK { return Err(err(self.ctx)); } }
        Ok(())
    }
    pub fn zero_grad(&mut self) { unsafe { vitrs_model_zero_grad(self.model); } }
}

/// `optimizer_step(model, lr)` (rv:949).
pub fn optimizer_step(model: &mut ViT, lr: f32) { unsafe { vitrs_model_optimizer_step(model.model, lr); } }

unsafe fn err(ctx: *mut vitrs_ctx) -> String { std::ffi::CStr::from_ptr(vitrs_last_error(ctx)).to_string_lossy().into_owned() }

impl Drop for ViT {
    fn drop(&mut self) { unsafe { vitrs_model_destroy(self.model); vitrs_ctx_destroy(self.ctx); } }
}
