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
