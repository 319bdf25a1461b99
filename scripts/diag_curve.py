"""Diagnostic (not a test): bf16 vs oracle loss-curve deviation by step range, and f32 per-tensor errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
from oracle import pyoracle as po
pkg = ge.load_package()
cfg = po.CONFIGS["tiny"]
for nb, lr in ((4, 1e-3), (8, 1e-3), (8, 5e-4), (16, 3e-4)):
    b, steps = 8, 100
    ref = po.ViT("tiny", seed=1337, init_mode=1)
    m = pkg.ViT("tiny", max_batch=b, mode=pkg.MODE_BF16, seed=1337, init_mode=1)
    rc, gc = [], []
    for step in range(steps):
        images, labels = po.synthetic_batch(cfg, b, step=step % nb)
        rc.append(ref.forward(images, labels)); ref.zero_grad(); ref.backward(); ref.update(lr)
        m.train_step(torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda(), lr)
        gc.append(m.mean_loss)
    rc, gc = np.array(rc), np.array(gc)
    d = np.abs(rc - gc)
    print(f"nb={nb} lr={lr}: ref first/last {rc[:4].mean():.3f}/{rc[-4:].mean():.3f} got last {gc[-4:].mean():.3f} | max dev by 10s:",
          " ".join(f"{d[i:i+10].max():.4f}" for i in range(0, 100, 10)), f"| mean {d.mean():.4f}")
    m.close()
for cfgname, b, im in (("tiny", 4, 0), ("tiny", 4, 1)):
    ref = po.ViT(cfgname, seed=1337, init_mode=im)
    m = pkg.ViT(cfgname, max_batch=b, mode=pkg.MODE_F32, seed=1337, init_mode=im)
    images, labels = po.synthetic_batch(cfg, b)
    ref.forward(images, labels); ref.zero_grad(); ref.backward()
    m.zero_grad(); m.forward(torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda()); m.backward()
    rel = lambda a, w: np.abs(a.astype(np.float64) - w).max() / max(np.abs(w).max(), 1e-30)
    print(cfgname, "init", im, "acts:", " ".join(f"{n}={rel(m.act(n).cpu().numpy(), ref.act(n)):.1e}" for n in po.ACT_NAMES))
    print(cfgname, "init", im, "grads:", " ".join(f"{n}={rel(m.grad(n).cpu().numpy(), ref.grad(n)):.1e}" for n in po.PARAM_NAMES))
    m.close()
