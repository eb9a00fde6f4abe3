"""CPU oracle for the dmip-b200 hot path — TEST INFRASTRUCTURE ONLY.

This package is a plain torch-CPU restatement (fp32 or fp64) of the algorithms
on the north-star hot path of maffos/Diffusion-Modelling-for-inverse-problems.
Every function cites the reference file:line it follows (paths relative to the
reference checkout).  It exists to *check* the CUDA path; it is never the thing
measured or shipped.

Import rules (enforced by tests/test_layout.py): only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import anything from `oracle/`.  The product package
(`diffusion-modelling-for-inverse-problems_b200/`, importable as `dmip`) never
does, and fails loudly when the CUDA library is missing.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md §4).  The oracle is therefore pinned against outputs of the
reference itself, produced in the authoring container by `oracle/make_golden.py`
(which imports the unmodified reference files from /root/reference through the
three import shims in `oracle/shims/`) and committed under `tests/golden/`.
`tests/test_oracle_golden.py` checks oracle == golden on every fixture.
Two boundaries stay "parity unpinned": the training-time sampler of `t`
(`sample_vp_truncated_q`, absent upstream — `t` is injected in every test) and
the VE-SDE / MMD items of BASELINE.json that have no reference counterpart.
"""

from . import vp, nets, sampler, losses, scatterometry, philox, metrics, mcmc  # noqa: F401
