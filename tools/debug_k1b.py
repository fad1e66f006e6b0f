"""Compare the K1b variants pairwise and run-to-run (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cases as C
from movedepth_b200 import ops
DEV = "cuda"
g = lambda t: t.to(DEV)
B, h, w, D = 2, 48, 160, 96
c = C.case_costvol("forward", B=B, h=h, w=w, D=D)
gv = g(torch.randn((B, 16, D, h, w), generator=torch.Generator().manual_seed(6)))
def run(flags, layout):
    ref, src = g(c["ref"]).requires_grad_(True), g(c["src"]).requires_grad_(True)
    got = ops.costvol_grouped(ref, src, g(c["K"]), g(c["invK"]), g(c["pose"][:, 0]), prior=g(c["prior"]), ratio=g(c["ratio"]), layout=layout, flags=flags)
    (got * gv).sum().backward()
    torch.cuda.synchronize()
    return ref.grad.clone(), src.grad.clone()
base = run(32, 0)
import collections
fails = collections.Counter()
for rep in range(100):
    for flags in (0, 0x400, 0x100, 0x200):
        gr, gs = run(flags, 1)
        dr = (gr - base[0]).abs()
        bad = (dr > 1e-3 * float(base[0].abs().max())).nonzero()
        if bad.shape[0]:
            fails[flags] += 1
            print("rep %d flags %d: %d bad, first %s" % (rep, flags, bad.shape[0], bad[0].tolist()), flush=True)
print("failures per variant over 100 reps:", dict(fails))
