"""The record loader (include/vitrs.h: vitrs_loader_*) on the committed CIFAR-10-layout fixture.

CPU part: record parsing, batch assembly by the native loader thread, epoch roll-over, shuffling as a permutation that is a
function of (seed, epoch) only.  GPU part: the tiny parity model trains from the loader (uint8 records, normalised on the device)
and its first step equals the oracle's on host-normalised pixels."""
import os

import numpy as np
import pytest

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cifar10_fixture.bin")
N, REC = 256, 3073


def records():
    raw = np.fromfile(FIXTURE, dtype=np.uint8).reshape(N, REC)
    return raw[:, 0].astype(np.int32), raw[:, 1:].reshape(N, 3, 32, 32)


def test_fixture_is_what_the_script_writes():
    import hashlib
    assert os.path.getsize(FIXTURE) == N * REC
    assert hashlib.md5(open(FIXTURE, "rb").read()).hexdigest() == "1be1bdb4e7c3ce8f61cd3a66f939b1be"


def test_sequential_batches_follow_the_file_and_wrap(vitrs):
    labels, images = records()
    ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=48, shuffle=False, drop_last=False, pinned=False)
    assert (ld.num_records, ld.batches_per_epoch, ld.num_classes_seen) == (256, 6, 10)
    at = 0
    for i in range(13):  # two epochs and one batch: 5 x 48 + 16, twice
        img, lab, epoch = ld.next()
        want = min(48, N - at)
        assert img.shape == (want, 3, 32, 32) and epoch == i // 6
        assert np.array_equal(lab, labels[at:at + want]) and np.array_equal(img, images[at:at + want])
        at = (at + want) % N
    ld.close()


def test_drop_last_and_two_files(vitrs):
    ld = vitrs.RecordLoader([FIXTURE, FIXTURE], image_size=32, batch=100, shuffle=False, drop_last=True, pinned=False)
    assert (ld.num_records, ld.batches_per_epoch) == (512, 5)
    labels, _ = records()
    both = np.concatenate([labels, labels])
    for i in range(7):
        img, lab, epoch = ld.next()
        assert len(lab) == 100 and epoch == i // 5
        assert np.array_equal(lab, both[(i % 5) * 100:(i % 5) * 100 + 100])
    ld.close()


def test_shuffle_is_a_permutation_per_epoch_and_deterministic(vitrs):
    labels, images = records()
    key = {images[i].tobytes(): i for i in range(N)}
    assert len(key) == N

    def two_epochs(seed):
        ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=32, shuffle=True, seed=seed, drop_last=True, pinned=False)
        order = []
        for _ in range(16):
            img, lab, epoch = ld.next()
            ids = [key[img[k].tobytes()] for k in range(len(lab))]
            assert np.array_equal(lab, labels[ids])  # labels travel with their images
            order.append((epoch, ids))
        ld.close()
        return [sum((ids for e, ids in order if e == ep), []) for ep in (0, 1)]

    a, b, c = two_epochs(7), two_epochs(7), two_epochs(8)
    assert a == b and a != c
    for ep in (0, 1):
        assert sorted(a[ep]) == list(range(N))
    assert a[0] != a[1] and a[0] != list(range(N))


def test_open_rejects_bad_files(vitrs, tmp_path):
    bad = tmp_path / "short.bin"
    bad.write_bytes(b"\x00" * 1000)
    with pytest.raises(vitrs.VitrsError):
        vitrs.RecordLoader(str(bad), image_size=32, batch=4, pinned=False)
    with pytest.raises(vitrs.VitrsError):
        vitrs.RecordLoader(str(tmp_path / "missing.bin"), image_size=32, batch=4, pinned=False)


@pytest.mark.gpu
def test_tiny_model_trains_from_the_loader_and_first_step_matches_oracle(vitrs):
    import torch
    from oracle import pyoracle as po
    labels, images = records()
    b = 32
    # first step, file order: the device normalises (x / 255 - 0.5) / 0.5; the oracle gets the same pixels normalised on the host
    ref = po.ViT("tiny", seed=1337, init_mode=1)
    host = ((images[:b].astype(np.float32) / 255.0) - 0.5) / 0.5
    ref_loss = ref.forward(host, labels[:b])
    m = vitrs.ViT("tiny", max_batch=b, mode=vitrs.MODE_F32, seed=1337, init_mode=1)
    ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=b, shuffle=False, ctx=m.ctx)
    loss, got_b = ld.train_step(m, 1e-3)
    assert got_b == b and abs(loss - ref_loss) <= 1e-4 * abs(ref_loss), (loss, ref_loss)
    ld.close(); m.close()
    # training: shuffled epochs in production mode, the loss falls
    m = vitrs.ViT("tiny", max_batch=b, mode=vitrs.MODE_BF16, seed=1337, init_mode=1)
    ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=b, shuffle=True, seed=3, ctx=m.ctx)
    curve = [ld.train_step(m, 2e-3)[0] for _ in range(15 * ld.batches_per_epoch)]
    first, last = np.mean(curve[:8]), np.mean(curve[-8:])
    assert np.isfinite(curve).all() and last < 0.6 * first, (first, last)
    ld.close(); m.close()


def test_sharded_loaders_partition_every_epoch(vitrs):
    """Data parallel: two ranks open the same file with the same seed; per epoch they see disjoint batches of the same (seed,
    epoch) order, the same number of batches, and together every record of the whole rounds exactly once."""
    labels, images = records()
    key = {images[i].tobytes(): i for i in range(N)}
    world, batch = 2, 24  # 256 // 24 = 10 batches -> 5 rounds of 2; the 16 left-over records are dropped this epoch
    seen = []
    for rank in range(world):
        ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=batch, shuffle=True, seed=11, pinned=False, rank=rank, world=world)
        assert ld.batches_per_epoch == 5
        per_epoch = {0: [], 1: []}
        for _ in range(10):
            img, lab, epoch = ld.next()
            per_epoch[epoch].append([key[img[k].tobytes()] for k in range(len(lab))])
        ld.close()
        assert [len(v) for v in per_epoch.values()] == [5, 5]
        seen.append(per_epoch)
    # the unsharded order of the same seed: rank r holds its batches r, r + 2, ...
    ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=batch, shuffle=True, seed=11, pinned=False)
    whole = []
    for _ in range(10):
        img, lab, epoch = ld.next()
        assert epoch == 0
        whole.append([key[img[k].tobytes()] for k in range(len(lab))])
    ld.close()
    for rank in range(world):
        assert seen[rank][0] == whole[rank::world]
    for ep in (0, 1):
        ids = sum(seen[0][ep] + seen[1][ep], [])
        assert len(ids) == len(set(ids)) == 240
    assert seen[0][0] != seen[0][1]
    with pytest.raises(vitrs.VitrsError):
        vitrs.RecordLoader(FIXTURE, image_size=32, batch=200, pinned=False, rank=0, world=2)  # one batch, two ranks


# ---- host-side transform: bilinear resize + the CIFAR augmentations, on several threads (vitrs_loader_open_transform) ----------


def test_resize_matches_torch_bilinear(vitrs):
    """32 x 32 records delivered at 224 x 224 (what a ViT-*/16 model takes): half-pixel-centre bilinear interpolation, as
    torch.nn.functional.interpolate(mode='bilinear', align_corners=False), rounded to uint8."""
    import torch
    labels, images = records()
    for out_size, workers in ((224, 4), (48, 1), (16, 2)):
        ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=32, shuffle=False, pinned=False, out_size=out_size, workers=workers)
        assert ld.image_size == out_size
        img, lab, _ = ld.next()
        assert img.shape == (32, 3, out_size, out_size) and np.array_equal(lab, labels[:32])
        want = torch.nn.functional.interpolate(torch.from_numpy(images[:32].astype(np.float32)), size=(out_size, out_size),
                                               mode="bilinear", align_corners=False).numpy()
        diff = np.abs(img.astype(np.float32) - want)
        assert diff.max() <= 0.5 + 1e-3, (out_size, diff.max())  # the rounding of an exact bilinear value
        ld.close()


def test_augmentation_is_flip_and_shift_of_the_record_and_independent_of_the_thread_count(vitrs):
    labels, images = records()
    pad = 4

    def epoch_images(workers, epochs=2, seed=5):
        ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=64, shuffle=False, seed=seed, pinned=False, workers=workers,
                                random_flip=True, crop_pad=pad)
        out = []
        for _ in range(epochs * 4):
            img, lab, ep = ld.next()
            out.append((ep, img.copy(), lab.copy()))
        ld.close()
        return out

    one, four = epoch_images(1), epoch_images(4)
    for (e1, i1, l1), (e4, i4, l4) in zip(one, four):
        assert e1 == e4 and np.array_equal(l1, l4) and np.array_equal(i1, i4)  # same bytes whatever the thread count
    assert not np.array_equal(one[0][1], one[4][1])  # another epoch, another draw
    assert not np.array_equal(one[0][1], epoch_images(1, epochs=1, seed=6)[0][1])
    # every delivered image is the record, mirrored or not, seen through a window shifted by at most `pad` pixels (zeros outside)
    padded = np.zeros((N, 3, 32 + 2 * pad, 32 + 2 * pad), np.uint8)
    padded[:, :, pad:pad + 32, pad:pad + 32] = images
    flips, shifts = 0, set()
    for b, (ep, img, lab) in enumerate(one[:4]):
        for k in range(0, 64, 7):
            rec = b * 64 + k
            found = None
            for flip in (0, 1):
                cand = img[k][:, :, ::-1] if flip else img[k]  # undo the mirror, then look for the window
                for dy in range(2 * pad + 1):
                    for dx in range(2 * pad + 1):
                        # a mirrored delivery is the mirror of a window of the record: mirror it back and compare with the window
                        # at the mirrored horizontal offset
                        ox = 2 * pad - dx if flip else dx
                        if np.array_equal(cand, padded[rec][:, dy:dy + 32, ox:ox + 32]):
                            found = (flip, dy, dx)
            assert found is not None, rec
            flips += found[0]
            shifts.add(found[1:])
    assert 0 < flips < 40 and len(shifts) > 10


def test_transform_loader_rejects_bad_options(vitrs):
    for kw in (dict(out_size=5000), dict(workers=1000), dict(crop_pad=40)):
        with pytest.raises(vitrs.VitrsError):
            vitrs.RecordLoader(FIXTURE, image_size=32, batch=8, pinned=False, **kw)


def test_resize_throughput_is_reported(vitrs, capsys):
    """Not a gate: prints what the host side delivers here (a ViT-B/16 step consumes ~8.8 k images/s per GPU)."""
    import time
    workers = min(8, os.cpu_count() or 1)
    ld = vitrs.RecordLoader(FIXTURE, image_size=32, batch=128, shuffle=True, pinned=False, out_size=224, workers=workers,
                            random_flip=True, crop_pad=4)
    ld.next()
    t0 = time.perf_counter()
    n = 0
    for _ in range(40):
        n += len(ld.next()[1])
    dt = time.perf_counter() - t0
    ld.close()
    with capsys.disabled():
        print(f"\n[loader] 32x32 -> 224x224 + flip + crop: {n / dt:.0f} images/s on {workers} threads")
    assert n == 40 * 128
