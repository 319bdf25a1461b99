"""Writes tests/golden/cifar10_fixture.bin: 256 records in the CIFAR-10 binary layout (1 label byte + 3072 uint8 samples,
channel-major 3x32x32; https://www.cs.toronto.edu/~kriz/cifar.html "binary version").  There is no network here, so the
pixels are synthetic but class-dependent (a per-class colour, stripe orientation and frequency, plus noise): a model can
learn them, which is what the loader test needs.  Deterministic: rerunning reproduces the committed file byte for byte."""
import os
import numpy as np

rng = np.random.RandomState(20261018)
n, size = 256, 32
yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
records = bytearray()
for i in range(n):
    label = i % 10
    colour = np.array([(label * 37) % 256, (label * 91 + 40) % 256, (label * 53 + 120) % 256], np.float32)
    angle = label * np.pi / 10.0
    wave = np.sin((xx * np.cos(angle) + yy * np.sin(angle)) * (0.3 + 0.08 * label) + rng.uniform(0, 2 * np.pi))
    img = colour[:, None, None] * 0.6 + 70.0 * wave[None] + rng.normal(0.0, 12.0, (3, size, size))
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    records += bytes([label]) + img.tobytes()
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cifar10_fixture.bin")
open(path, "wb").write(bytes(records))
print(path, len(records), "bytes")
