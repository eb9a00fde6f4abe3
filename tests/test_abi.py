"""C-ABI checks that need no GPU: libdmip_sm100.so loads, exports every symbol include/dmip.h declares, keeps the
struct layouts the ctypes bindings assume, and refuses to compute without an sm_100 device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmip.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"#if defined\(DMIP_DEBUG\).*?#endif", "", src, flags=re.S)      # debug-build-only hooks
    return sorted(set(re.findall(r"\b(dmip_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import dmip
    if not os.path.exists(dmip.library_path()):
        dmip.build()
    return dmip._lib.lib()


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("dmip_version", "dmip_last_error", "dmip_pack_mlp", "dmip_sampler_em_vp", "dmip_mlp_forward",
              "dmip_loss_fwd_bwd", "dmip_posterior_loss_fwd_bwd", "dmip_surrogate_score",
              "dmip_sampler_workspace_bytes", "dmip_loss_workspace_bytes"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    import dmip
    out = subprocess.run(["nm", "-D", "--defined-only", dmip.library_path()], capture_output=True, text=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    for s in declared_symbols():
        getattr(lib, s)                        # dlsym
    # probes, self-tests and micro-benchmarks are not part of the product library (they live in tools/probe)
    assert not [s for s in exported if s.startswith("dmip_debug")], "debug entry points leaked into the product"


def test_version_and_error_string(lib):
    assert lib.dmip_version() == 100
    assert isinstance(lib.dmip_last_error(), bytes)


def _sizeof_from_c(struct_names):
    """Compile a tiny C program against include/dmip.h and read back sizeof() of each struct."""
    import tempfile
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in struct_names)
    src = f'#include <stdio.h>\n#include "dmip.h"\nint main(void){{{body}return 0;}}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    return {l.split()[0]: int(l.split()[1]) for l in out.splitlines()}


def test_ctypes_struct_layouts_match_the_header():
    """The header is plain C (compiles with gcc -std=c99) and the Python mirrors have identical sizes."""
    from dmip import _lib, losses, mcmc, metrics, nets, posterior, utils_scatterometry
    mirrors = {"DmipMlpGrad": nets.DmipMlpGrad, "DmipMlp": _lib.DmipMlp, "DmipSampler": _lib.DmipSampler, "DmipForward": _lib.DmipForward,
               "DmipLoss": losses.DmipLoss, "DmipPosteriorLoss": posterior.DmipPosteriorLoss,
               "DmipSurrogate": utils_scatterometry.DmipSurrogate, "DmipHistogram": metrics.DmipHistogram,
               "DmipMetropolis": mcmc.DmipMetropolis}
    sizes = _sizeof_from_c(list(mirrors))
    for name, cls in mirrors.items():
        assert C.sizeof(cls) == sizes[name], (name, C.sizeof(cls), sizes[name])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU refusal path")
def test_no_cpu_fallback(lib):
    import dmip
    from dmip import _lib
    from dmip.models.diffusion import CDE
    assert not dmip.is_available()
    assert lib.dmip_device_ok() == 0
    m = CDE(2, 2, [16, 16])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2), num_samples=4, num_steps=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sde.a(torch.zeros(4, 2), torch.zeros(4, 2), torch.zeros(4, 1))
    with pytest.raises(RuntimeError):
        _lib.require_gpu()
    # the C entry points themselves refuse (DMIP_EARCH), they do not compute on the host
    d = _lib.DmipSampler()
    d.variant, d.precision, d.xdim, d.ydim, d.n_obs, d.n_per_obs, d.num_steps = 0, 0, 2, 2, 1, 4, 2
    assert lib.dmip_sampler_em_vp(C.byref(d), None) == -2
    assert b"no CPU" in lib.dmip_last_error()
