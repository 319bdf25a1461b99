"""ctypes view of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see vit_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.  Function names and argument order are the reference's
(train_vit.rs:376-670); arrays are float32 / int32 numpy, C-contiguous.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

PARAM_NAMES = ["patchw", "patchb", "cls", "wpe", "ln1w", "ln1b", "qkvw", "qkvb", "attprojw",
               "attprojb", "ln2w", "ln2b", "fcw", "fcb", "fcprojw", "fcprojb", "lnfw", "lnfb",
               "headw", "headb"]
ACT_NAMES = ["encoded", "ln1", "ln1_mean", "ln1_rstd", "qkv", "atty", "preatt", "att", "attproj",
             "residual2", "ln2", "ln2_mean", "ln2_rstd", "fch", "fch_gelu", "fcproj", "residual3",
             "lnf", "lnf_mean", "lnf_rstd", "logits", "probs", "losses"]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only, no reference sources)."""
    src = [os.path.join(_HERE, f) for f in ("vit_oracle.c", "vit_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class ViTConfig(C.Structure):
    _fields_ = [("max_seq_len", C.c_int), ("vocab_size", C.c_int), ("num_layers", C.c_int),
                ("num_heads", C.c_int), ("channels", C.c_int), ("image_size", C.c_int),
                ("patch_size", C.c_int), ("num_classes", C.c_int), ("causal", C.c_int)]


_lib = None
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.vit_build.restype = C.c_void_p
        _lib.vit_build.argtypes = [C.POINTER(ViTConfig), C.c_uint64, C.c_int]
        for name in ("vit_param_ptr", "vit_grad_ptr", "vit_act_ptr", "vit_grad_act_ptr"):
            getattr(_lib, name).restype = _f32p
            getattr(_lib, name).argtypes = [C.c_void_p, C.c_int]
        _lib.vit_act_size.restype = C.c_size_t
        _lib.vit_act_size.argtypes = [C.c_void_p, C.c_int]
        _lib.vit_param_sizes.argtypes = [C.POINTER(ViTConfig), C.POINTER(C.c_size_t)]
        _lib.vit_forward.argtypes = [C.c_void_p, _f32p, _i32p, C.c_int]
        _lib.vit_backward.argtypes = [C.c_void_p]
        _lib.vit_zero_grad.argtypes = [C.c_void_p]
        _lib.vit_free.argtypes = [C.c_void_p]
        _lib.vit_update.argtypes = [C.c_void_p] + [C.c_float] * 5
        _lib.vit_rand_u01.restype = C.c_float
        _lib.vit_rand_u01.argtypes = [C.c_uint64] * 3
        _lib.vit_fill_uniform.argtypes = [_f32p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_float, C.c_float]
        _lib.adamw_step.argtypes = [_f32p, _f32p, _f32p, _f32p, C.c_size_t] + [C.c_float] * 5 + [C.c_int]
        _lib.sgd_step.argtypes = [_f32p, _f32p, C.c_size_t, C.c_float]
        _lib.vit_oracle_num_threads.restype = C.c_int
        _lib.vit_mean_loss.restype = C.c_float
        _lib.vit_mean_loss.argtypes = [C.c_void_p]
        _lib.vit_set_dloss_scale.argtypes = [C.c_void_p, C.c_float]
    return _lib


def _p(a):
    if a is None:
        return None
    if a.dtype == np.float32:
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(_f32p)
    if a.dtype == np.int32:
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(_i32p)
    raise TypeError(a.dtype)


def _call(name, *args):
    fn = getattr(lib(), name)
    fn.restype = None
    conv = []
    for a in args:
        if isinstance(a, np.ndarray) or a is None:
            conv.append(_p(a))
        elif isinstance(a, float):
            conv.append(C.c_float(a))
        else:
            conv.append(C.c_int(int(a)))
    fn(*conv)


# ---- reference-named operators ----------------------------------------------------------
def residual_forward(out, inp1, inp2, n): _call("residual_forward", out, inp1, inp2, n)
def residual_backward(dinp1, dinp2, dout, n): _call("residual_backward", dinp1, dinp2, dout, n)
def matmul_forward(out, inp, weight, bias, b, t, c, oc): _call("matmul_forward", out, inp, weight, bias, b, t, c, oc)
def matmul_backward(dinp, dweight, dbias, dout, inp, weight, b, t, c, oc):
    _call("matmul_backward", dinp, dweight, dbias, dout, inp, weight, b, t, c, oc)
def attention_forward(out, preatt, att, inp, b, t, c, nh, causal=1):
    _call("attention_forward_ex", out, preatt, att, inp, b, t, c, nh, causal)
def attention_backward(dinp, dpreatt, datt, dout, inp, att, b, t, c, nh, causal=1):
    _call("attention_backward_ex", dinp, dpreatt, datt, dout, inp, att, b, t, c, nh, causal)
def layernorm_forward(out, mean, rstd, inp, weight, bias, b, t, c):
    _call("layernorm_forward", out, mean, rstd, inp, weight, bias, b, t, c)
def layernorm_backward(dinp, dweight, dbias, dout, inp, weight, mean, rstd, b, t, c):
    _call("layernorm_backward", dinp, dweight, dbias, dout, inp, weight, mean, rstd, b, t, c)
def gelu_forward(out, inp, n): _call("gelu_forward", out, inp, n)
def gelu_backward(dinp, inp, dout, n): _call("gelu_backward", dinp, inp, dout, n)
def softmax_forward(probs, logits, b, t, v): _call("softmax_forward", probs, logits, b, t, v)
def crossentropy_forward(losses, probs, targets, b, t, v): _call("crossentropy_forward", losses, probs, targets, b, t, v)
def crossentropy_softmax_backward(dlogits, dlosses, probs, targets, b, t, v):
    _call("crossentropy_softmax_backward", dlogits, dlosses, probs, targets, b, t, v)
def encoder_forward(encoded, inputs, wte, wpe, b, t, c): _call("encoder_forward", encoded, inputs, wte, wpe, b, t, c)
def encoder_backward(dwte, dwpe, dencoded, inputs, b, t, c): _call("encoder_backward", dwte, dwpe, dencoded, inputs, b, t, c)
def patch_embed_forward(encoded, images, patchw, patchb, cls, wpe, b, img, patch, c):
    _call("patch_embed_forward", encoded, images, patchw, patchb, cls, wpe, b, img, patch, c)
def patch_embed_backward(dpatchw, dpatchb, dcls, dwpe, dencoded, images, b, img, patch, c):
    _call("patch_embed_backward", dpatchw, dpatchb, dcls, dwpe, dencoded, images, b, img, patch, c)


def adamw_step(params, grads, m, v, lr, beta1, beta2, eps, weight_decay, step):
    lib().adamw_step(_p(params), _p(grads), _p(m), _p(v), params.size, lr, beta1, beta2, eps, weight_decay, step)


def sgd_step(params, grads, lr):
    lib().sgd_step(_p(params), _p(grads), params.size, lr)


def rand_u01(seed, stream, n):
    """numpy restatement of vit_rand_u01 (checked against the C in tests)."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        x = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
             + np.uint64(stream) * np.uint64(0xD1B54A32D192ED03) + idx)
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return (x >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def fill_uniform(n, seed, stream, lo, hi):
    return (np.float32(lo) + (np.float32(hi) - np.float32(lo)) * rand_u01(seed, stream, n)).astype(np.float32)


def synthetic_batch(cfg, b, seed=1337, step=0):
    """images U[-1,1) [B,3,H,W] fp32 and labels uniform in [0,classes)."""
    n = b * 3 * cfg["image_size"] * cfg["image_size"]
    images = fill_uniform(n, seed, 1000 + 2 * step, -1.0, 1.0).reshape(b, 3, cfg["image_size"], cfg["image_size"])
    labels = np.minimum((rand_u01(seed, 1001 + 2 * step, b) * cfg["num_classes"]).astype(np.int32),
                        cfg["num_classes"] - 1).astype(np.int32)
    return images, labels


CONFIGS = {
    # BASELINE.json configs[0] — build-defined tiny model
    "tiny": dict(image_size=32, patch_size=4, channels=64, num_layers=2, num_heads=4, num_classes=10),
    "ti16": dict(image_size=224, patch_size=16, channels=192, num_layers=12, num_heads=3, num_classes=1000),
    "s16": dict(image_size=224, patch_size=16, channels=384, num_layers=12, num_heads=6, num_classes=1000),
    "b16": dict(image_size=224, patch_size=16, channels=768, num_layers=12, num_heads=12, num_classes=1000),
    "b8": dict(image_size=224, patch_size=8, channels=768, num_layers=12, num_heads=12, num_classes=1000),
}


def make_config(name_or_dict, causal=0):
    d = dict(CONFIGS[name_or_dict]) if isinstance(name_or_dict, str) else dict(name_or_dict)
    t = (d["image_size"] // d["patch_size"]) ** 2 + 1
    return ViTConfig(t, d["num_classes"], d["num_layers"], d["num_heads"], d["channels"],
                     d["image_size"], d["patch_size"], d["num_classes"], causal)


class ViT:
    """Mirror of the reference's `ViT` (rusty_vit.rs:63-450) over the C oracle."""

    def __init__(self, cfg, seed=1337, init_mode=0, causal=0):
        self.cfg = make_config(cfg, causal) if not isinstance(cfg, ViTConfig) else cfg
        self._h = lib().vit_build(C.byref(self.cfg), seed, init_mode)
        sizes = (C.c_size_t * 20)()
        lib().vit_param_sizes(C.byref(self.cfg), sizes)
        self.param_sizes = list(sizes)
        self.num_parameters = sum(self.param_sizes)
        self.batch_size = 0
        self._keep = None

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:  # (module globals are already gone at interpreter shutdown)
            lib().vit_free(self._h)
            self._h = None

    def _view(self, getter, idx, n):
        ptr = getattr(lib(), getter)(self._h, idx)
        return np.ctypeslib.as_array(ptr, shape=(n,))

    def param(self, name):
        i = PARAM_NAMES.index(name)
        return self._view("vit_param_ptr", i, self.param_sizes[i])

    def grad(self, name):
        i = PARAM_NAMES.index(name)
        return self._view("vit_grad_ptr", i, self.param_sizes[i])

    def params_flat(self):
        return self._view("vit_param_ptr", 0, self.num_parameters)

    def grads_flat(self):
        return self._view("vit_grad_ptr", 0, self.num_parameters)

    def act(self, name):
        i = ACT_NAMES.index(name)
        return self._view("vit_act_ptr", i, lib().vit_act_size(self._h, i))

    def grad_act(self, name):
        i = ACT_NAMES.index(name)
        return self._view("vit_grad_act_ptr", i, lib().vit_act_size(self._h, i))

    @property
    def mean_loss(self):
        return float(lib().vit_mean_loss(self._h))

    def set_dloss_scale(self, s):
        lib().vit_set_dloss_scale(self._h, s)

    def forward(self, images, targets, b=None):
        images = np.ascontiguousarray(images, dtype=np.float32)
        b = images.shape[0] if b is None else b
        tg = None if targets is None else np.ascontiguousarray(targets, dtype=np.int32)
        self._keep = (images, tg)
        self.batch_size = b
        lib().vit_forward(self._h, _p(images), _p(tg), b)
        return self.mean_loss

    def zero_grad(self):
        lib().vit_zero_grad(self._h)

    def backward(self):
        lib().vit_backward(self._h)

    def update(self, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01):
        lib().vit_update(self._h, lr, beta1, beta2, eps, weight_decay)
