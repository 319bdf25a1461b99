"""vitrs_b200 — Python host-side mirror of the ViT.rs model / operator surface over libvitrs.so.

The product is the C-ABI library `libvitrs.so` (include/vitrs.h, sources in csrc/): hand-written
sm_100a CUDA behind the reference's llm.c-style operator signatures (train_vit.rs:376-670) and
its `ViT` model object (rusty_vit.rs:63-450).  This module only binds it with ctypes and uses
PyTorch for device memory and streams — there is no Python or CPU compute path, and importing
the ops without the built library raises.

The directory is named after the reference (`vit.rs_b200`), which is not an importable module
name; load it with `__graft_entry__.load_package()` (registers it as `vitrs_b200`).
"""
import ctypes as C
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VITRS_LIB: A/B aid for kernel experiments (scripts/ab_bench.sh); the product is the in-tree library
LIB_PATH = os.environ.get("VITRS_LIB") or os.path.join(_HERE, "libvitrs.so")

MODE_F32, MODE_BF16 = 0, 1

PARAM_NAMES = ["patchw", "patchb", "cls", "wpe", "ln1w", "ln1b", "qkvw", "qkvb", "attprojw",
               "attprojb", "ln2w", "ln2b", "fcw", "fcb", "fcprojw", "fcprojb", "lnfw", "lnfb",
               "headw", "headb"]
ACT_NAMES = ["encoded", "ln1", "ln1_mean", "ln1_rstd", "qkv", "atty", "preatt", "att", "attproj",
             "residual2", "ln2", "ln2_mean", "ln2_rstd", "fch", "fch_gelu", "fcproj", "residual3",
             "lnf", "lnf_mean", "lnf_rstd", "logits", "probs", "losses"]

CONFIGS = {
    # BASELINE.json configs[0]: build-defined tiny model
    "tiny": dict(image_size=32, patch_size=4, channels=64, num_layers=2, num_heads=4, num_classes=10),
    "ti16": dict(image_size=224, patch_size=16, channels=192, num_layers=12, num_heads=3, num_classes=1000),
    "s16": dict(image_size=224, patch_size=16, channels=384, num_layers=12, num_heads=6, num_classes=1000),
    "b16": dict(image_size=224, patch_size=16, channels=768, num_layers=12, num_heads=12, num_classes=1000),
    "b8": dict(image_size=224, patch_size=8, channels=768, num_layers=12, num_heads=12, num_classes=1000),
}


class VitrsError(RuntimeError):
    pass


class Config(C.Structure):
    """vitrs_config — ViTConfig (rusty_vit.rs:10-16) plus the ViT fields."""
    _fields_ = [("max_seq_len", C.c_int), ("vocab_size", C.c_int), ("num_layers", C.c_int),
                ("num_heads", C.c_int), ("channels", C.c_int), ("image_size", C.c_int),
                ("patch_size", C.c_int), ("num_classes", C.c_int), ("causal", C.c_int)]


def make_config(name_or_dict, causal=0):
    d = dict(CONFIGS[name_or_dict]) if isinstance(name_or_dict, str) else dict(name_or_dict)
    t = (d["image_size"] // d["patch_size"]) ** 2 + 1
    return Config(t, d["num_classes"], d["num_layers"], d["num_heads"], d["channels"],
                  d["image_size"], d["patch_size"], d["num_classes"], causal)


def train_flops_per_image(cfg):
    """Algorithmic training flops per image: 3 x forward matmul flops."""
    d = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    p, c, l, v = d["patch_size"], d["channels"], d["num_layers"], d["num_classes"]
    n = (d["image_size"] // p) ** 2
    t = n + 1
    fwd = 2 * n * (3 * p * p) * c + l * (24 * t * c * c + 4 * t * t * c) + 2 * c * v
    return 3 * fwd


class Footprint(C.Structure):
    """vitrs_footprint — device bytes of a model by what they hold (vitrs_model_footprint)."""
    _fields_ = [(n, C.c_uint64) for n in ("num_parameters", "weights_f32", "grads_f32", "weights_bf16", "adam_moments",
                                          "zero1_master_shard", "exchange_buffer", "activations", "activation_grads",
                                          "workspace", "staging", "total")] + [("train_flops_per_image", C.c_double)]


class GemmPlan(C.Structure):
    """vitrs_gemm_plan_t — what a GEMM call of given extents launches (vitrs_gemm_plan)."""
    _fields_ = [(n, C.c_int) for n in ("kernel", "tile_m", "tile_n", "cta_group", "stages", "splits", "k_blocks_per_split",
                                       "tiles", "grid")]


class LoaderOptions(C.Structure):
    """vitrs_loader_options — record loader with host-side augmentation and resize (vitrs_loader_open_transform)."""
    _fields_ = [("image_size", C.c_int), ("out_size", C.c_int), ("label_bytes", C.c_int), ("batch", C.c_int), ("shuffle", C.c_int),
                ("drop_last", C.c_int), ("seed", C.c_uint64), ("rank", C.c_int), ("world", C.c_int), ("workers", C.c_int),
                ("random_flip", C.c_int), ("crop_pad", C.c_int)]


PLAN_SIMT, PLAN_TCGEN05 = 0, 1
PLAN_NO_SMALL, PLAN_SINGLE_CTA, PLAN_PATCH_TC = 1, 2, 4
EPI_NONE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_GELU_BWD, EPI_ACCUM_F32, EPI_PATCH, EPI_ROWDOT, EPI_BIAS_GELU_ONLY = range(9)

_vp, _f32p, _i32p, _u16p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
_int, _sz, _f, _u64 = C.c_int, C.c_size_t, C.c_float, C.c_uint64

# every symbol include/vitrs.h declares: name -> (restype, argtypes)
_SIGNATURES = {
    "vitrs_ctx_create": (_int, [C.POINTER(C.c_void_p), _int]),
    "vitrs_ctx_destroy": (_int, [_vp]),
    "vitrs_ctx_set_stream": (_int, [_vp, _vp]),
    "vitrs_ctx_reset_stream": (_int, [_vp]),
    "vitrs_ctx_stream": (_vp, [_vp]),
    "vitrs_ctx_synchronize": (_int, [_vp]),
    "vitrs_last_error": (C.c_char_p, [_vp]),
    "vitrs_launch_count": (_u64, [_vp]),
    "vitrs_version": (C.c_char_p, []),
    "vitrs_profile_begin": (_int, [_vp]),
    "vitrs_profile_end": (_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_int)]),
    "vitrs_malloc": (_int, [_vp, C.POINTER(C.c_void_p), _sz]),
    "vitrs_free": (_int, [_vp, _vp]),
    "vitrs_malloc_host": (_int, [_vp, C.POINTER(C.c_void_p), _sz]),
    "vitrs_free_host": (_int, [_vp, _vp]),
    "vitrs_memcpy_h2d": (_int, [_vp, _vp, _vp, _sz]),
    "vitrs_memcpy_d2h": (_int, [_vp, _vp, _vp, _sz]),
    "vitrs_memset": (_int, [_vp, _vp, _int, _sz]),
    "vitrs_cast_f32_to_bf16": (_int, [_vp, _vp, _vp, _sz]),
    "vitrs_cast_bf16_to_f32": (_int, [_vp, _vp, _vp, _sz]),
    "vitrs_gemm_bf16": (_int, [_vp, _vp, _vp, _vp] + [_int] * 9),
    "vitrs_gemm_bf16_fused": (_int, [_vp] * 8 + [_int] * 9),
    "vitrs_sgd_step": (_int, [_vp, _vp, _vp, _sz, _f, _vp]),
    "vitrs_adamw_step": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _int, _vp]),
    "vitrs_fill_uniform": (_int, [_vp, _vp, _sz, _u64, _u64, _f, _f]),
    "vitrs_crossentropy_forward_f32": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int]),
    "vitrs_crossentropy_softmax_backward_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int]),
    "vitrs_softmax_forward_f32": (_int, [_vp, _vp, _vp, _int, _int, _int]),
    "vitrs_encoder_forward_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int]),
    "vitrs_encoder_backward_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int]),
    "vitrs_patch_embed_forward_f32": (_int, [_vp] + [_vp] * 6 + [_int] * 4),
    "vitrs_patch_embed_backward_f32": (_int, [_vp] + [_vp] * 6 + [_int] * 4),
    "vitrs_attention_forward_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int]),
    "vitrs_attention_backward_f32": (_int, [_vp] + [_vp] * 6 + [_int] * 5),
    "vitrs_attention_forward_bf16": (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int]),
    "vitrs_attention_backward_bf16": (_int, [_vp] + [_vp] * 5 + [_int] * 5),
    "vitrs_model_create": (_int, [_vp, C.POINTER(Config), _int, _int, C.POINTER(C.c_void_p)]),
    "vitrs_model_destroy": (_int, [_vp]),
    "vitrs_model_init_parameters": (_int, [_vp, _u64, _int]),
    "vitrs_model_save_checkpoint": (_int, [_vp, C.c_char_p]),
    "vitrs_model_load_checkpoint": (_int, [_vp, C.c_char_p]),
    "vitrs_model_num_parameters": (_sz, [_vp]),
    "vitrs_model_sync_parameters": (_int, [_vp]),
    "vitrs_model_param_view": (_int, [_vp, _int, _int, C.POINTER(C.c_void_p), C.POINTER(_sz)]),
    "vitrs_model_act_view": (_int, [_vp, _int, _int, C.POINTER(C.c_void_p), C.POINTER(_sz), C.POINTER(_int)]),
    "vitrs_model_set_dloss_scale": (_int, [_vp, _f]),
    "vitrs_model_forward": (_int, [_vp, _vp, _vp, _int]),
    "vitrs_model_zero_grad": (_int, [_vp]),
    "vitrs_model_backward": (_int, [_vp]),
    "vitrs_model_optimizer_step": (_int, [_vp, _f]),
    "vitrs_model_update": (_int, [_vp, _f, _f, _f, _f, _f]),
    "vitrs_model_mean_loss": (_int, [_vp, C.POINTER(_f)]),
    "vitrs_model_prefetch_host": (_int, [_vp, _vp, _vp, _int]),
    "vitrs_model_train_step_host": (_int, [_vp, _vp, _vp, _int, _f, _f, _f, _f, _f, C.POINTER(_f)]),
    "vitrs_model_train_step": (_int, [_vp, _vp, _vp, _int, _f, _f, _f, _f, _f]),
    "vitrs_model_step_graph_replays": (_int, [_vp, C.POINTER(_u64)]),
    "vitrs_model_set_input_norm": (_int, [_vp, C.POINTER(_f), C.POINTER(_f)]),
    "vitrs_model_forward_u8": (_int, [_vp, _vp, _int, _vp, _int]),
    "vitrs_model_train_step_u8": (_int, [_vp, _vp, _int, _vp, _int, _f, _f, _f, _f, _f]),
    "vitrs_model_prefetch_host_u8": (_int, [_vp, _vp, _vp, _int]),
    "vitrs_model_train_step_host_u8": (_int, [_vp, _vp, _int, _vp, _int, _f, _f, _f, _f, _f, C.POINTER(_f)]),
    "vitrs_comm_unique_id": (_int, [_vp, _vp]),
    "vitrs_comm_init": (_int, [_vp, _vp, _int, _int]),
    "vitrs_comm_init_config": (_int, [_vp, _vp, _int, _int, _int]),
    "vitrs_comm_async_error": (_int, [_vp, C.POINTER(_int)]),
    "vitrs_comm_destroy": (_int, [_vp]),
    "vitrs_ctx_error_flags": (_int, [_vp, C.POINTER(_int)]),
    "vitrs_model_set_comm_dtype": (_int, [_vp, _int]),
    "vitrs_model_enable_zero1": (_int, [_vp]),
    "vitrs_model_gather_parameters": (_int, [_vp]),
    "vitrs_model_optimizer_state_bytes": (_int, [_vp, C.POINTER(_sz)]),
    "vitrs_zero_partition": (_int, [C.POINTER(Config), _int, _int, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "vitrs_model_footprint": (_int, [C.POINTER(Config), _int, _int, _int, _int, C.POINTER(Footprint)]),
    "vitrs_infer_footprint": (_int, [C.POINTER(Config), _int, C.POINTER(_u64), C.POINTER(_u64)]),
    "vitrs_gemm_plan": (_int, [_int] * 8 + [C.POINTER(GemmPlan)]),
    "vitrs_comm_world": (_int, [_vp, C.POINTER(_int), C.POINTER(_int)]),
    "vitrs_model_allreduce_grads": (_int, [_vp]),
    "vitrs_grad_bucket": (_int, [C.POINTER(Config), _int, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_int), C.POINTER(_int)]),
    "vitrs_allreduce_f32": (_int, [_vp, _vp, _sz]),
    "vitrs_loader_open": (_int, [_vp, C.POINTER(C.c_char_p), _int, _int, _int, _int, _int, _u64, _int, C.POINTER(C.c_void_p)]),
    "vitrs_loader_open_sharded": (_int, [_vp, C.POINTER(C.c_char_p), _int, _int, _int, _int, _int, _u64, _int, _int, _int, C.POINTER(C.c_void_p)]),
    "vitrs_loader_open_transform": (_int, [_vp, C.POINTER(C.c_char_p), _int, C.POINTER(LoaderOptions), C.POINTER(C.c_void_p)]),
    "vitrs_loader_image_size": (_int, [_vp]),
    "vitrs_loader_close": (_int, [_vp]),
    "vitrs_loader_info": (_int, [_vp, C.POINTER(_sz), C.POINTER(_int), C.POINTER(_int)]),
    "vitrs_loader_next": (_int, [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(_int), C.POINTER(_u64)]),
    "vitrs_model_train_step_loader": (_int, [_vp, _vp, _f, _f, _f, _f, _f, C.POINTER(_f), C.POINTER(_int)]),
    "vitrs_infer_create": (_int, [_vp, _int, C.POINTER(C.c_void_p)]),
    "vitrs_infer_destroy": (_int, [_vp]),
    "vitrs_infer_forward": (_int, [_vp, _vp, _int]),
    "vitrs_infer_forward_u8": (_int, [_vp, _vp, _int, _int]),
    "vitrs_infer_outputs": (_int, [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "vitrs_infer_forward_host": (_int, [_vp, _vp, _int, _vp]),
    "vitrs_infer_forward_host_u8": (_int, [_vp, _vp, _int, _int, _vp]),
    "vitrs_infer_set_graph": (_int, [_vp, _int]),
    "vitrs_infer_stats": (_int, [_vp, C.POINTER(_sz), C.POINTER(_u64)]),
}
# the two-mode operator families share argument lists
for _m in ("f32", "bf16"):
    _SIGNATURES.update({
        f"vitrs_residual_forward_{_m}": (_int, [_vp, _vp, _vp, _vp, _int]),
        f"vitrs_matmul_forward_{_m}": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int]),
        f"vitrs_layernorm_forward_{_m}": (_int, [_vp] + [_vp] * 6 + [_int] * 3),
        f"vitrs_gelu_forward_{_m}": (_int, [_vp, _vp, _vp, _int]),
        f"vitrs_residual_backward_{_m}": (_int, [_vp, _vp, _vp, _vp, _int]),
        f"vitrs_matmul_backward_{_m}": (_int, [_vp] + [_vp] * 6 + [_int] * 4),
        f"vitrs_layernorm_backward_{_m}": (_int, [_vp] + [_vp] * 8 + [_int] * 3),
        f"vitrs_gelu_backward_{_m}": (_int, [_vp, _vp, _vp, _vp, _int]),
    })

EXPORTED_SYMBOLS = sorted(_SIGNATURES)

_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VitrsError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(make -C vit.rs_b200/csrc). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def _ptr(x):
    """Device (torch) or host (numpy) buffer -> raw address; None -> NULL."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    assert x.is_contiguous(), "tensors crossing the C ABI must be contiguous"
    return x.data_ptr()


class Context:
    """vitrs_ctx: device, streams, workspace, optional NCCL communicator."""

    def __init__(self, device=0, use_torch_stream=True):
        self._h = C.c_void_p()
        rc = lib().vitrs_ctx_create(C.byref(self._h), device)
        if rc != 0:
            raise VitrsError(f"vitrs_ctx_create({device}) failed ({rc}): {lib().vitrs_last_error(None).decode()}")
        self.device = device
        if use_torch_stream:
            import torch
            self.set_stream(torch.cuda.current_stream(device).cuda_stream)

    def check(self, rc):
        if rc != 0:
            raise VitrsError(f"libvitrs error {rc}: {lib().vitrs_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream):
        self.check(lib().vitrs_ctx_set_stream(self._h, cuda_stream))

    def synchronize(self):
        self.check(lib().vitrs_ctx_synchronize(self._h))

    @property
    def launches(self):
        return int(lib().vitrs_launch_count(self._h))

    def close(self):
        if self._h:
            lib().vitrs_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def profile_begin(self):
        """Bracket every tcgen05 GEMM launch with CUDA events on the launching stream until profile_end()."""
        self.check(lib().vitrs_profile_begin(self._h))

    def profile_end(self):
        """-> (total GEMM milliseconds, total GEMM flops, launches) since profile_begin(); synchronises."""
        ms, fl, n = C.c_double(), C.c_double(), C.c_int()
        self.check(lib().vitrs_profile_end(self._h, C.byref(ms), C.byref(fl), C.byref(n)))
        return ms.value, fl.value, n.value

    # ---- data parallel ------------------------------------------------------------------
    def comm_unique_id(self):
        buf = (C.c_char * 128)()
        self.check(lib().vitrs_comm_unique_id(self._h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def comm_init(self, uid, rank, world, max_ctas=None):
        """max_ctas: cap on the thread blocks NCCL may occupy per collective (None: the library default, 8; 0: NCCL's own)."""
        buf = C.create_string_buffer(uid, 128)
        if max_ctas is None:
            self.check(lib().vitrs_comm_init(self._h, C.cast(buf, C.c_void_p), rank, world))
        else:
            self.check(lib().vitrs_comm_init_config(self._h, C.cast(buf, C.c_void_p), rank, world, max_ctas))

    def comm_async_error(self):
        """ncclCommGetAsyncError: 0 when healthy; raises VitrsError otherwise."""
        r = C.c_int()
        self.check(lib().vitrs_comm_async_error(self._h, C.byref(r)))
        return r.value

    def error_flags(self):
        f = C.c_int()
        self.check(lib().vitrs_ctx_error_flags(self._h, C.byref(f)))
        return f.value

    def allreduce(self, t):
        self.check(lib().vitrs_allreduce_f32(self._h, _ptr(t), t.numel()))


def model_footprint(cfg, max_batch, mode=MODE_BF16, world=1, zero1=False):
    """Device bytes of ViT(cfg, max_batch, mode) by what they hold, and the algorithmic flops per image (host arithmetic
    only: the sizing functions vitrs_model_create itself uses).  -> dict"""
    c = make_config(cfg) if not isinstance(cfg, Config) else cfg
    f = Footprint()
    rc = lib().vitrs_model_footprint(C.byref(c), max_batch, mode, world, int(bool(zero1)), C.byref(f))
    if rc != 0:
        raise VitrsError(f"vitrs_model_footprint failed ({rc})")
    return {n: getattr(f, n) for n, _ in Footprint._fields_}


def max_batch_for(cfg, hbm_bytes=180 * 10 ** 9, mode=MODE_BF16, world=1, zero1=False, reserve=4 * 10 ** 9):
    """Largest per-GPU batch whose footprint fits `hbm_bytes` minus `reserve` (CUDA context, NCCL, allocator slack)."""
    lo, hi = 0, 1
    fits = lambda b: model_footprint(cfg, b, mode, world, zero1)["total"] <= hbm_bytes - reserve
    while fits(hi) and hi < 1 << 24:
        lo, hi = hi, hi * 2
    while hi - lo > 1:
        mid = (lo + hi) // 2
        lo, hi = (mid, hi) if fits(mid) else (lo, mid)
    return lo


def infer_footprint(cfg, max_batch):
    """-> (workspace bytes, staging bytes) of InferenceEngine(model of cfg, max_batch); host arithmetic only."""
    c = make_config(cfg) if not isinstance(cfg, Config) else cfg
    ws, st = C.c_uint64(), C.c_uint64()
    rc = lib().vitrs_infer_footprint(C.byref(c), max_batch, C.byref(ws), C.byref(st))
    if rc != 0:
        raise VitrsError(f"vitrs_infer_footprint failed ({rc})")
    return ws.value, st.value


def gemm_plan(M, N, K, a_mn=False, b_mn=False, epilogue=EPI_NONE, sm_count=148, flags=0):
    """What the library launches for a dense [M, K] x [N, K] bf16 GEMM with this epilogue on a device with sm_count SMs
    (host arithmetic only: the routing function gemm_tc_bf16 itself uses).  -> dict"""
    p = GemmPlan()
    rc = lib().vitrs_gemm_plan(M, N, K, int(a_mn), int(b_mn), epilogue, sm_count, flags, C.byref(p))
    if rc != 0:
        raise VitrsError(f"vitrs_gemm_plan failed ({rc})")
    d = {n: getattr(p, n) for n, _ in GemmPlan._fields_}
    d["kernel"] = "tcgen05" if p.kernel == PLAN_TCGEN05 else "simt"
    return d


def step_gemms(cfg, batch):
    """The bf16 GEMM calls of one production training step of `cfg` at `batch` images, in the order of model.cu
    (forward_bf16 / backward_bf16): list of (name, M, N, K, a_mn, b_mn, epilogue).  The class head is fp32 SIMT and not listed."""
    d = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    c, p = d["channels"], d["patch_size"]
    t = (d["image_size"] // p) ** 2 + 1
    rows, kdim = batch * t, 3 * p * p
    fwd = [("patch", rows, c, kdim, 0, 0, EPI_PATCH)]
    blk_f = [("qkv", rows, 3 * c, c, 0, 0, EPI_BIAS), ("attproj", rows, c, c, 0, 0, EPI_BIAS_RESIDUAL),
             ("fc", rows, 4 * c, c, 0, 0, EPI_BIAS_GELU), ("fcproj", rows, c, 4 * c, 0, 0, EPI_BIAS_RESIDUAL)]
    blk_b = [("fcproj_dx", rows, 4 * c, c, 0, 1, EPI_GELU_BWD), ("fcproj_dw", c, 4 * c, rows, 1, 1, EPI_ACCUM_F32),
             ("fc_dx", rows, c, 4 * c, 0, 1, EPI_NONE), ("fc_dw", 4 * c, c, rows, 1, 1, EPI_ACCUM_F32),
             ("attproj_dx", rows, c, c, 0, 1, EPI_ROWDOT if c == 64 * d["num_heads"] else EPI_NONE), ("attproj_dw", c, c, rows, 1, 1, EPI_ACCUM_F32),
             ("qkv_dx", rows, c, 3 * c, 0, 1, EPI_NONE), ("qkv_dw", 3 * c, c, rows, 1, 1, EPI_ACCUM_F32)]
    return fwd + blk_f * d["num_layers"] + blk_b * d["num_layers"] + [("patch_dw", c, kdim, rows, 1, 1, EPI_ACCUM_F32)]


def zero_partition(cfg, world):
    """The ZeRO-1 partition (host arithmetic only): per bucket (z_off, z_len, z_big, shard)."""
    c = make_config(cfg) if not isinstance(cfg, Config) else cfg
    out = []
    for b in range(c.num_layers + 2):
        zo, zl, zb, sh = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        rc = lib().vitrs_zero_partition(C.byref(c), world, b, C.byref(zo), C.byref(zl), C.byref(zb), C.byref(sh))
        if rc != 0:
            raise VitrsError(f"vitrs_zero_partition failed ({rc})")
        out.append((zo.value, zl.value, zb.value, sh.value))
    return out


def grad_buckets(cfg, with_kind=False):
    """The gradient-exchange schedule (host arithmetic only): list of buckets, each a list of (offset, count)
    — or (offset, count, big) with with_kind: big = a GEMM weight matrix (sharded by ZeRO-1)."""
    c = make_config(cfg) if not isinstance(cfg, Config) else cfg
    out = []
    for b in range(c.num_layers + 2):
        off, cnt, big, n = (C.c_size_t * 12)(), (C.c_size_t * 12)(), (C.c_int * 12)(), C.c_int()
        rc = lib().vitrs_grad_bucket(C.byref(c), b, off, cnt, big, C.byref(n))
        if rc != 0:
            raise VitrsError(f"vitrs_grad_bucket failed ({rc})")
        out.append([(off[i], cnt[i], big[i]) if with_kind else (off[i], cnt[i]) for i in range(n.value)])
    return out


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---- L1 operators: the reference's names and argument order (train_vit.rs:376-670) ----------
def _sfx(t):
    import torch
    if t.dtype == torch.float32:
        return "f32"
    if t.dtype == torch.bfloat16:
        return "bf16"
    raise TypeError(f"unsupported dtype {t.dtype}")


def _call(ctx, name, *args):
    ctx = ctx or default_context()
    conv = [_ptr(a) if not isinstance(a, (int, float)) or a is None else a for a in args]
    ctx.check(getattr(lib(), name)(ctx._h, *conv))


def residual_forward(out, inp1, inp2, n, ctx=None):
    _call(ctx, f"vitrs_residual_forward_{_sfx(out)}", out, inp1, inp2, n)


def residual_backward(dinp1, dinp2, dout, n, ctx=None):
    _call(ctx, f"vitrs_residual_backward_{_sfx(dout)}", dinp1, dinp2, dout, n)


def matmul_forward(out, inp, weight, bias, b, t, c, oc, ctx=None):
    _call(ctx, f"vitrs_matmul_forward_{_sfx(out)}", out, inp, weight, bias, b, t, c, oc)


def matmul_backward(dinp, dweight, dbias, dout, inp, weight, b, t, c, oc, ctx=None):
    _call(ctx, f"vitrs_matmul_backward_{_sfx(dout)}", dinp, dweight, dbias, dout, inp, weight, b, t, c, oc)


def attention_forward(out, preatt, att, inp, b, t, c, nh, causal=1, ctx=None):
    """fp32: reference signature (preatt/att may be None).  bf16: pass lse as `preatt`, att=None."""
    if _sfx(out) == "f32":
        _call(ctx, "vitrs_attention_forward_f32", out, preatt, att, inp, b, t, c, nh, causal)
    else:
        _call(ctx, "vitrs_attention_forward_bf16", out, preatt, inp, b, t, c, nh, causal)


def attention_backward(dinp, dpreatt, datt, dout, inp, att, b, t, c, nh, causal=1, ctx=None):
    _call(ctx, "vitrs_attention_backward_f32", dinp, dpreatt, datt, dout, inp, att, b, t, c, nh, causal)


def attention_backward_bf16(dinp, dout, out, lse, inp, b, t, c, nh, causal=0, ctx=None):
    _call(ctx, "vitrs_attention_backward_bf16", dinp, dout, out, lse, inp, b, t, c, nh, causal)


def layernorm_forward(out, mean, rstd, inp, weight, bias, b, t, c, ctx=None):
    _call(ctx, f"vitrs_layernorm_forward_{_sfx(out)}", out, mean, rstd, inp, weight, bias, b, t, c)


def layernorm_backward(dinp, dweight, dbias, dout, inp, weight, mean, rstd, b, t, c, ctx=None):
    _call(ctx, f"vitrs_layernorm_backward_{_sfx(dout)}", dinp, dweight, dbias, dout, inp, weight, mean, rstd, b, t, c)


def gelu_forward(out, inp, n, ctx=None):
    _call(ctx, f"vitrs_gelu_forward_{_sfx(out)}", out, inp, n)


def gelu_backward(dinp, inp, dout, n, ctx=None):
    _call(ctx, f"vitrs_gelu_backward_{_sfx(dout)}", dinp, inp, dout, n)


def softmax_forward(probs, logits, b, t, v, ctx=None):
    _call(ctx, "vitrs_softmax_forward_f32", probs, logits, b, t, v)


def crossentropy_forward(losses, probs, targets, b, t, v, ctx=None):
    _call(ctx, "vitrs_crossentropy_forward_f32", losses, probs, targets, b, t, v)


def crossentropy_softmax_backward(dlogits, dlosses, probs, targets, b, t, v, ctx=None):
    _call(ctx, "vitrs_crossentropy_softmax_backward_f32", dlogits, dlosses, probs, targets, b, t, v)


def encoder_forward(encoded, inputs, wte, wpe, b, t, c, ctx=None):
    _call(ctx, "vitrs_encoder_forward_f32", encoded, inputs, wte, wpe, b, t, c)


def encoder_backward(dwte, dwpe, dencoded, inputs, b, t, c, ctx=None):
    _call(ctx, "vitrs_encoder_backward_f32", dwte, dwpe, dencoded, inputs, b, t, c)


def patch_embed_forward(encoded, images, patchw, patchb, cls, wpe, b, img, patch, c, ctx=None):
    _call(ctx, "vitrs_patch_embed_forward_f32", encoded, images, patchw, patchb, cls, wpe, b, img, patch, c)


def patch_embed_backward(dpatchw, dpatchb, dcls, dwpe, dencoded, images, b, img, patch, c, ctx=None):
    _call(ctx, "vitrs_patch_embed_backward_f32", dpatchw, dpatchb, dcls, dwpe, dencoded, images, b, img, patch, c)


EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_GELU_BWD, EPI_BIAS_GELU_ONLY = 1, 2, 3, 4, 8


def gemm_bf16_fused(D, D2, aux, bias, a_colsum, A, B, M, N, K, lda, ldb, ldd, a_mn_major=0, b_mn_major=0, epilogue=EPI_BIAS, ctx=None):
    """The tcgen05 GEMM with one of the training step's fused epilogues (include/vitrs.h: vitrs_gemm_bf16_fused)."""
    _call(ctx, "vitrs_gemm_bf16_fused", D, D2, aux, bias, a_colsum, A, B, M, N, K, lda, ldb, ldd, a_mn_major, b_mn_major, epilogue)


def gemm_bf16(D, A, B, M, N, K, lda, ldb, ldd, a_mn_major=0, b_mn_major=0, out_f32_accumulate=0, ctx=None):
    _call(ctx, "vitrs_gemm_bf16", D, A, B, M, N, K, lda, ldb, ldd, a_mn_major, b_mn_major, out_f32_accumulate)


def adamw_step(params, grads, m, v, lr, beta1, beta2, eps, weight_decay, step, shadow=None, ctx=None):
    ctx = ctx or default_context()
    ctx.check(lib().vitrs_adamw_step(ctx._h, _ptr(params), _ptr(grads), _ptr(m), _ptr(v), params.numel(), lr, beta1,
                                     beta2, eps, weight_decay, step, _ptr(shadow)))


def sgd_step(params, grads, lr, shadow=None, ctx=None):
    ctx = ctx or default_context()
    ctx.check(lib().vitrs_sgd_step(ctx._h, _ptr(params), _ptr(grads), params.numel(), lr, _ptr(shadow)))


def fill_uniform(dst, seed, stream, lo, hi, ctx=None):
    ctx = ctx or default_context()
    ctx.check(lib().vitrs_fill_uniform(ctx._h, _ptr(dst), dst.numel(), seed, stream, lo, hi))


class _DeviceView:
    """__cuda_array_interface__ holder so torch can alias library-owned device memory."""

    def __init__(self, ptr, count, typestr, owner):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}
        self._owner = owner


def _tensor_view(ptr, count, elem, device, owner):
    import torch
    if not ptr or count == 0:
        return None
    if elem == 4:
        return torch.as_tensor(_DeviceView(ptr, count, "<f4", owner), device=f"cuda:{device}")
    t = torch.as_tensor(_DeviceView(ptr, count, "<i2", owner), device=f"cuda:{device}")
    return t.view(torch.bfloat16)


class ViT:
    """`struct ViT` + `impl ViT` of the reference (rusty_vit.rs:63-450) on the GPU.

    forward(images, targets, b) / backward() / optimizer_step(lr) keep the reference's names
    and meaning; update(lr, ...) is the AdamW form of the step (DEVIATIONS D8).  Parameters,
    gradients and activations are exposed as the reference's named views (torch tensors that
    alias the library's flat device buffers).
    """

    def __init__(self, cfg, max_batch, mode=MODE_BF16, seed=1337, init_mode=0, causal=0, ctx=None, init=True):
        self.ctx = ctx or default_context()
        self.cfg_dict = dict(CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
        self.cfg = make_config(self.cfg_dict, causal)
        self.mode = mode
        self.max_batch = max_batch
        self._h = C.c_void_p()
        self.ctx.check(lib().vitrs_model_create(self.ctx._h, C.byref(self.cfg), max_batch, mode, C.byref(self._h)))
        self.num_parameters = int(lib().vitrs_model_num_parameters(self._h))
        self.batch_size = 0
        if init:
            self.init_parameters(seed, init_mode)

    @classmethod
    def build_from_checkpoint(cls, path, max_batch, mode=MODE_BF16, ctx=None):
        """ViT::build_from_checkpoint (rusty_vit.rs:79-259): config from the 256-int header, then the parameters."""
        with open(path, "rb") as f:
            header = struct.unpack("<256i", f.read(1024))
        cfg = dict(image_size=header[7], patch_size=header[8], channels=header[6], num_layers=header[4],
                   num_heads=header[5], num_classes=header[9])
        m = cls(cfg, max_batch, mode, causal=header[10], ctx=ctx, init=False)
        m.load_checkpoint(path)
        return m

    def close(self):
        if self._h:
            lib().vitrs_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_parameters(self, seed=1337, init_mode=0):
        self.ctx.check(lib().vitrs_model_init_parameters(self._h, seed, init_mode))

    def save_checkpoint(self, path):
        self.ctx.check(lib().vitrs_model_save_checkpoint(self._h, path.encode()))

    def load_checkpoint(self, path):
        self.ctx.check(lib().vitrs_model_load_checkpoint(self._h, path.encode()))

    def sync_parameters(self):
        self.ctx.check(lib().vitrs_model_sync_parameters(self._h))

    def _pview(self, which, name):
        ptr, cnt = C.c_void_p(), C.c_size_t()
        self.ctx.check(lib().vitrs_model_param_view(self._h, which, PARAM_NAMES.index(name), C.byref(ptr), C.byref(cnt)))
        return _tensor_view(ptr.value, cnt.value, 4, self.ctx.device, self)

    def param(self, name): return self._pview(0, name)
    def grad(self, name): return self._pview(1, name)
    def adam_m(self, name): return self._pview(2, name)
    def adam_v(self, name): return self._pview(3, name)

    def _flat(self, which):
        ptr, cnt = C.c_void_p(), C.c_size_t()
        self.ctx.check(lib().vitrs_model_param_view(self._h, which, 0, C.byref(ptr), C.byref(cnt)))
        return _tensor_view(ptr.value, self.num_parameters, 4, self.ctx.device, self)

    def params_flat(self): return self._flat(0)
    def grads_flat(self): return self._flat(1)

    def _aview(self, which, name):
        ptr, cnt, elem = C.c_void_p(), C.c_size_t(), C.c_int()
        self.ctx.check(lib().vitrs_model_act_view(self._h, which, ACT_NAMES.index(name), C.byref(ptr), C.byref(cnt), C.byref(elem)))
        return _tensor_view(ptr.value, cnt.value, elem.value, self.ctx.device, self)

    def act(self, name): return self._aview(0, name)
    def grad_act(self, name): return self._aview(1, name)

    def set_comm_dtype(self, dtype):
        """Gradient exchange payload: "bf16" (default, one packed message per bucket) or "f32" (exact in-place slices)."""
        self.ctx.check(lib().vitrs_model_set_comm_dtype(self._h, {"f32": 0, "bf16": 1}.get(dtype, dtype)))

    def enable_zero1(self):
        """ZeRO-1: shard fp32 master weights and AdamW moments 1/world per bucket (include/vitrs.h)."""
        self.ctx.check(lib().vitrs_model_enable_zero1(self._h))

    def gather_parameters(self):
        self.ctx.check(lib().vitrs_model_gather_parameters(self._h))

    @property
    def optimizer_state_bytes(self):
        n = C.c_size_t()
        self.ctx.check(lib().vitrs_model_optimizer_state_bytes(self._h, C.byref(n)))
        return n.value

    @property
    def step_graph_replays(self):
        n = C.c_uint64()
        self.ctx.check(lib().vitrs_model_step_graph_replays(self._h, C.byref(n)))
        return n.value

    def set_dloss_scale(self, s):
        self.ctx.check(lib().vitrs_model_set_dloss_scale(self._h, s))

    def forward(self, images, targets, b=None):
        """images: cuda float32 [b,3,H,W]; targets: cuda int32 [b] or None (logits only, mean_loss = -1)."""
        b = images.shape[0] if b is None else b
        self._keep = (images, targets)
        self.batch_size = b
        self.ctx.check(lib().vitrs_model_forward(self._h, _ptr(images), _ptr(targets), b))

    def zero_grad(self):
        self.ctx.check(lib().vitrs_model_zero_grad(self._h))

    def backward(self):
        self.ctx.check(lib().vitrs_model_backward(self._h))

    def optimizer_step(self, lr):
        """optimizer_step(model, lr) — the reference's SGD (rusty_vit.rs:949-955)."""
        self.ctx.check(lib().vitrs_model_optimizer_step(self._h, lr))

    def update(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
        self.ctx.check(lib().vitrs_model_update(self._h, lr, beta1, beta2, eps, weight_decay))

    @property
    def mean_loss(self):
        out = C.c_float()
        self.ctx.check(lib().vitrs_model_mean_loss(self._h, C.byref(out)))
        return out.value

    def train_step(self, images, targets, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
        b = images.shape[0]
        self._keep = (images, targets)
        self.batch_size = b
        self.ctx.check(lib().vitrs_model_train_step(self._h, _ptr(images), _ptr(targets), b, lr, beta1, beta2, eps, weight_decay))

    # ---- raw uint8 image batches (include/vitrs.h: the data path in front of the step) ----
    NCHW, NHWC = 0, 1

    def set_input_norm(self, mean, std):
        """Per-channel (x / 255 - mean) / std applied inside the patch embedding's im2col pass."""
        m3, s3 = (_f * 3)(*[float(x) for x in mean]), (_f * 3)(*[float(x) for x in std])
        self.ctx.check(lib().vitrs_model_set_input_norm(self._h, m3, s3))

    def forward_u8(self, images, targets, layout=0):
        """images: cuda uint8 [b,3,H,W] (NCHW) or [b,H,W,3] (NHWC)."""
        b = images.shape[0]
        self._keep = (images, targets)
        self.batch_size = b
        self.ctx.check(lib().vitrs_model_forward_u8(self._h, _ptr(images), layout, _ptr(targets), b))

    def train_step_u8(self, images, targets, lr, layout=0, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
        b = images.shape[0]
        self._keep = (images, targets)
        self.batch_size = b
        self.ctx.check(lib().vitrs_model_train_step_u8(self._h, _ptr(images), layout, _ptr(targets), b, lr, beta1, beta2, eps, weight_decay))

    def prefetch_host_u8(self, h_images, h_labels):
        self.ctx.check(lib().vitrs_model_prefetch_host_u8(self._h, _ptr(h_images), _ptr(h_labels), h_images.shape[0]))

    def train_step_host_u8(self, h_images, h_labels, lr, layout=0, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
        """One step from HOST uint8 buffers: a quarter of the fp32 bytes over PCIe, normalised on the device."""
        out = C.c_float()
        b = h_images.shape[0]
        self.batch_size = b
        self.ctx.check(lib().vitrs_model_train_step_host_u8(self._h, _ptr(h_images), layout, _ptr(h_labels), b, lr, beta1, beta2, eps,
                                                            weight_decay, C.byref(out)))
        return out.value

    def prefetch_host(self, h_images, h_labels):
        self.ctx.check(lib().vitrs_model_prefetch_host(self._h, _ptr(h_images), _ptr(h_labels), h_images.shape[0]))

    def train_step_host(self, h_images, h_labels, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
        """One step from HOST buffers (pinned torch tensors or numpy): H2D, step, D2H of the loss."""
        out = C.c_float()
        b = h_images.shape[0]
        self.batch_size = b
        self.ctx.check(lib().vitrs_model_train_step_host(self._h, _ptr(h_images), _ptr(h_labels), b, lr, beta1, beta2, eps,
                                                         weight_decay, C.byref(out)))
        return out.value


class InferenceEngine:
    """ViT::forward without targets (rusty_vit.rs:339-350) as a serving object: borrows `model`'s parameters, owns a ping-pong
    workspace (no [L, ...] activation arena) and replays the forward as a CUDA graph (include/vitrs.h: vitrs_infer_*)."""

    def __init__(self, model, max_batch, use_graph=True):
        self.model = model  # keeps the parameters alive
        self.ctx = model.ctx
        self.max_batch = max_batch
        self.num_classes = model.cfg_dict["num_classes"]
        self._h = C.c_void_p()
        self.ctx.check(lib().vitrs_infer_create(model._h, max_batch, C.byref(self._h)))
        if not use_graph:
            self.ctx.check(lib().vitrs_infer_set_graph(self._h, 0))

    def close(self):
        if self._h:
            lib().vitrs_infer_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _outputs(self, b):
        lg, pr = C.c_void_p(), C.c_void_p()
        self.ctx.check(lib().vitrs_infer_outputs(self._h, C.byref(lg), C.byref(pr)))
        n = b * self.num_classes
        return (_tensor_view(lg.value, n, 4, self.ctx.device, self).view(b, self.num_classes),
                _tensor_view(pr.value, n, 4, self.ctx.device, self).view(b, self.num_classes))

    def forward(self, images):
        """images: cuda fp32 [b,3,H,W]; returns (logits, probs) views [b, classes] (valid until the next call)."""
        b = images.shape[0]
        self._keep = images
        self.ctx.check(lib().vitrs_infer_forward(self._h, _ptr(images), b))
        return self._outputs(b)

    def forward_u8(self, images, layout=0):
        b = images.shape[0]
        self._keep = images
        self.ctx.check(lib().vitrs_infer_forward_u8(self._h, _ptr(images), layout, b))
        return self._outputs(b)

    def forward_host(self, h_images, h_logits=None):
        """Host images (numpy / pinned torch, fp32 NCHW) -> host logits [b, classes]: H2D, graph replay, D2H, synchronised."""
        import numpy as np
        b = h_images.shape[0]
        out = h_logits if h_logits is not None else np.empty((b, self.num_classes), dtype=np.float32)
        self.ctx.check(lib().vitrs_infer_forward_host(self._h, _ptr(h_images), b, _ptr(out)))
        return out

    def forward_host_u8(self, h_images, layout=0, h_logits=None):
        import numpy as np
        b = h_images.shape[0]
        out = h_logits if h_logits is not None else np.empty((b, self.num_classes), dtype=np.float32)
        self.ctx.check(lib().vitrs_infer_forward_host_u8(self._h, _ptr(h_images), layout, b, _ptr(out)))
        return out

    def stats(self):
        ws, rep = C.c_size_t(), C.c_uint64()
        self.ctx.check(lib().vitrs_infer_stats(self._h, C.byref(ws), C.byref(rep)))
        return {"workspace_bytes": ws.value, "graph_replays": rep.value}


class RecordLoader:
    """CIFAR-layout record files -> shuffled uint8 batches assembled by a native loader thread into pinned host slots
    (include/vitrs.h: vitrs_loader_*).  ctx=None uses pageable memory (no GPU needed)."""

    def __init__(self, paths, image_size, batch, label_bytes=1, shuffle=True, seed=0, drop_last=True, ctx=None, pinned=True, rank=0, world=1,
                 out_size=0, workers=0, random_flip=False, crop_pad=0):
        """out_size / random_flip / crop_pad / workers: host-side bilinear resize and the CIFAR augmentations, spread over
        `workers` threads (vitrs_loader_open_transform); all left at their defaults the records are delivered as stored."""
        paths = [paths] if isinstance(paths, str) else list(paths)
        self.ctx = ctx if ctx is not None else (default_context() if pinned else None)
        self.batch = batch
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        self._h = C.c_void_p()
        h = self.ctx._h if self.ctx else None
        if out_size or workers or random_flip or crop_pad:
            opt = LoaderOptions(image_size, out_size, label_bytes, batch, int(shuffle), int(drop_last), seed, rank, world, workers,
                                int(random_flip), crop_pad)
            rc = lib().vitrs_loader_open_transform(h, arr, len(paths), C.byref(opt), C.byref(self._h))
        else:
            rc = lib().vitrs_loader_open_sharded(h, arr, len(paths), image_size, label_bytes, batch, int(shuffle), seed, int(drop_last),
                                                 rank, world, C.byref(self._h))
        if rc != 0:
            raise VitrsError(f"vitrs_loader_open failed ({rc})" + (f": {lib().vitrs_last_error(self.ctx._h).decode()}" if self.ctx else ""))
        n, bpe, ncls = C.c_size_t(), C.c_int(), C.c_int()
        lib().vitrs_loader_info(self._h, C.byref(n), C.byref(bpe), C.byref(ncls))
        self.num_records, self.batches_per_epoch, self.num_classes_seen = n.value, bpe.value, ncls.value
        self.image_size = lib().vitrs_loader_image_size(self._h)  # side of the delivered images

    def next(self):
        """(images uint8 [b,3,H,W], labels int32 [b], epoch) as numpy views of the loader's slot: valid until the call after the next."""
        pi, pl, b, ep = C.c_void_p(), C.c_void_p(), C.c_int(), C.c_uint64()
        rc = lib().vitrs_loader_next(self._h, C.byref(pi), C.byref(pl), C.byref(b), C.byref(ep))
        if rc != 0:
            raise VitrsError(f"vitrs_loader_next failed ({rc})")
        n = b.value
        img = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_uint8)), shape=(n, 3, self.image_size, self.image_size))
        lab = np.ctypeslib.as_array(C.cast(pl, C.POINTER(C.c_int32)), shape=(n,))
        return img, lab, ep.value

    def train_step(self, model, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
        """One AdamW step of `model` on the next batch (vitrs_model_train_step_loader); returns (loss, batch size)."""
        loss, b = C.c_float(), C.c_int()
        model.ctx.check(lib().vitrs_model_train_step_loader(model._h, self._h, lr, beta1, beta2, eps, weight_decay, C.byref(loss), C.byref(b)))
        model.batch_size = b.value
        return loss.value, b.value

    def close(self):
        if self._h:
            lib().vitrs_loader_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
