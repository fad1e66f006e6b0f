"""CPU oracle for the MOVEDepth dense hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (CPU, fp32/fp64) restatement of the reference
algorithm for the path named in BASELINE.json (`north_star`).  It exists so that
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` leg have something to check the CUDA path against (and to
time on the host cores).  Nothing under `movedepth_b200/` may import it: the
product path runs hand-written sm_100a kernels through the C-ABI library and
fails loudly when that library is missing.

Pinning status: the reference ships no tests or golden vectors,
so the pins are outputs of the reference code itself, generated in the build
container by `tests/golden/make_golden.py` (which imports /root/reference) and
committed under `tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks every
function here against those vectors.

The arithmetic of the reference lives in a third-party dependency that is not
vendored: PyTorch (pinned torch==1.7.1 / torchvision==0.8.2 in the reference's
environment.yml:14-15).  The oracle therefore calls the same torch primitives
(`F.grid_sample`, `F.avg_pool2d`, `F.unfold`, convolutions, batch-norm) on CPU;
what is restated here is the reference's own composition of them.
"""
