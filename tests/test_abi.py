"""CPU-side checks of the drop-in boundary: libvitrs.so loads, exports every symbol that
include/vitrs.h declares (and nothing the header does not know), and refuses to run without a
GPU instead of falling back to a CPU path.  No compute calls are made here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vitrs.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vitrs_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_table_agree(vitrs):
    assert declared_symbols() == vitrs.EXPORTED_SYMBOLS


def test_library_exports_every_declared_symbol(vitrs):
    assert os.path.exists(vitrs.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    handle = ctypes.CDLL(vitrs.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(handle, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", vitrs.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (vitrs_[a-z0-9_]+)$", out, flags=re.M)))
    extra = [s for s in exported if s not in declared_symbols()]
    assert extra == [], extra


def test_rust_crate_declares_every_symbol():
    """rust/vitrs-sys/src/ffi.rs is generated from the header (scripts/gen_rust_ffi.py): it must be current and complete."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "scripts", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    want = gen.render(gen.parse_header(HEADER))
    have = open(os.path.join(ROOT, "rust", "vitrs-sys", "src", "ffi.rs")).read()
    assert have == want, "stale: python scripts/gen_rust_ffi.py > rust/vitrs-sys/src/ffi.rs"
    assert sorted(re.findall(r"pub fn (vitrs_[a-z0-9_]+)\(", have)) == declared_symbols()
    # the wrapper only calls what ffi.rs declares
    wrapper = open(os.path.join(ROOT, "rust", "vitrs-sys", "src", "lib.rs")).read()
    used = set(re.findall(r"\b(vitrs_[a-z0-9_]+)\(", wrapper))
    assert used <= set(declared_symbols()), sorted(used - set(declared_symbols()))


def test_library_has_no_oracle_or_torch_dependency(vitrs):
    out = subprocess.run(["ldd", vitrs.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out and "libnccl" not in out  # NCCL is dlopen'ed at comm_init


def test_sass_is_blackwell_native(vitrs):
    """tcgen05.mma / tcgen05.ld / TMA must be in the shipped SASS (B200_PROFILING.md mnemonics)."""
    out = subprocess.run(["cuobjdump", "-sass", vitrs.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    sass = out.stdout
    assert "sm_100a" in sass
    assert re.search(r"\bUTC\w*MMA\b", sass), "no tcgen05.mma in SASS"
    assert "LDTM" in sass and "UTMALDG" in sass


def test_no_cpu_fallback_without_gpu(vitrs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vitrs.VitrsError):
        vitrs.Context(0)
    # ops refuse a NULL context instead of computing anything
    assert vitrs.lib().vitrs_residual_forward_f32(None, None, None, None, 4) == -2


def test_config_table_and_flops(vitrs):
    # train GFLOP / image
    for name, want in (("ti16", 7.522), ("s16", 27.593), ("b16", 105.383), ("b8", 468.890)):
        assert abs(vitrs.train_flops_per_image(name) / 1e9 - want) < 2e-3, name
    cfg = vitrs.make_config("b16")
    assert (cfg.max_seq_len, cfg.channels, cfg.num_heads, cfg.num_layers) == (197, 768, 12, 12)


def test_header_is_plain_c_and_links_from_c(vitrs, tmp_path):
    """The boundary is a C ABI: include/vitrs.h compiles as strict C99 (no C++ or torch type in any signature) and a C program
    links libvitrs.so, calls the host-only planning entry points and is refused a context when there is no GPU."""
    import shutil
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HEADER],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = tmp_path / "cabi.c"
    src.write_text(r'''
#include <stdio.h>
#include "vitrs.h"
int main(void) {
    vitrs_config cfg = {0, 0, 12, 12, 768, 224, 16, 1000, 0};
    vitrs_footprint f;
    vitrs_gemm_plan_t p;
    vitrs_ctx* ctx = NULL;
    if (vitrs_model_footprint(&cfg, 1024, VITRS_MODE_BF16, 1, 0, &f) != VITRS_OK) return 1;
    if (vitrs_gemm_plan(201728, 2304, 768, 0, 0, 1, 148, 0, &p) != VITRS_OK) return 2;
    printf("%s|%llu|%d|%d|%d|%d\n", vitrs_version(), (unsigned long long)f.num_parameters, p.kernel, p.tile_n, p.grid,
           vitrs_ctx_create(&ctx, 0));
    return 0;
}
''')
    exe = tmp_path / "cabi"
    libdir = os.path.dirname(vitrs.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe), "-L", libdir, "-lvitrs",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    version, params, kernel, tile_n, grid, ctx_rc = out.stdout.strip().split("|")
    assert version.startswith("vitrs-b200") and int(params) == 86567656 and (int(kernel), int(tile_n), int(grid)) == (1, 256, 148)
    import torch
    if not torch.cuda.is_available():
        assert int(ctx_rc) < 0  # no device: an error code, never a CPU path
