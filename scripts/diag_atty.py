"""Diagnostic: per-layer error of atty / qkv / residual2 against the oracle for one bf16 forward of a model config."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
from oracle import pyoracle as po
pkg = ge.load_package()
name, b = os.environ.get("CFG", "ti16"), int(os.environ.get("B", 4))
cfg = po.CONFIGS[name]
ref = po.ViT(name, seed=1337, init_mode=1)
images, labels = po.synthetic_batch(cfg, b)
ref.forward(images, labels)
m = pkg.ViT(name, max_batch=b, mode=pkg.MODE_BF16, seed=1337, init_mode=1)
for rep in range(3):
    m.forward(torch.from_numpy(images).cuda(), torch.from_numpy(labels).cuda())
    L = cfg["num_layers"]
    line = []
    for act in ("qkv", "atty", "residual2"):
        got = m.act(act).float().cpu().numpy().reshape(L, -1)
        want = ref.act(act)[:got.size].reshape(L, -1)
        line.append(act + " " + " ".join(f"{np.abs(got[l]-want[l]).max()/np.abs(want[l]).max():.1e}" for l in range(L)))
    print(f"rep {rep}:", " | ".join(line))
