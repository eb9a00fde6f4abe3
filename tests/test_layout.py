"""Repository rules that the judge checks mechanically: the oracle is test infrastructure only, the product never
imports it, nothing run on the GPU box reads /root/reference, and no compatibility layers are on the hot path."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "diffusion-modelling-for-inverse-problems_b200")


def _py_files(top):
    for d, _, fs in os.walk(top):
        if "__pycache__" in d:
            continue
        for f in fs:
            if f.endswith(".py"):
                yield os.path.join(d, f)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    for top in (PKG, os.path.join(ROOT, "dmip")):
        for f in _py_files(top):
            assert not pat.search(open(f).read()), f"{f} imports oracle/"


def test_bench_only_uses_the_oracle_in_the_cpu_legs():
    src = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle|import oracle", src)]
    assert uses, "bench.py must time the oracle port as cpu_baseline"
    for pos in uses:
        fn = re.findall(r"^def (\w+)", src[:pos], flags=re.M)[-1]
        assert fn.startswith("cpu_"), f"oracle imported in {fn}()"


def test_gpu_side_code_does_not_read_the_reference_checkout():
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    files += list(_py_files(PKG)) + [f for f in _py_files(os.path.join(ROOT, "tests"))]
    for f in files:
        if os.path.basename(f) == "test_layout.py":
            continue
        assert "/root/reference" not in open(f).read(), f


def test_no_compatibility_layers_on_the_hot_path():
    for f in _py_files(PKG):
        src = open(f).read()
        assert not re.search(r"^\s*import triton|torch\.compile\(|^\s*import tilelang", src, re.M), f


def test_oracle_header_says_test_infrastructure():
    for f in _py_files(os.path.join(ROOT, "oracle")):
        if os.sep + "shims" + os.sep in f or os.path.basename(f) == "make_golden.py":
            continue
        assert "TEST INFRASTRUCTURE ONLY" in open(f).read(), f


def test_required_documents_exist():
    for name in ("DESIGN.md", "INTEGRATION.md", "include/dmip.h", "bench.py", "__graft_entry__.py"):
        assert os.path.exists(os.path.join(ROOT, name)), name
