"""CPU: the vectorised DataLoader against a literal per-frame / per-object restatement of the reference's
algorithm (utils/data_loader.py:66-151, 185-247) on a real SDD extract (tests/golden/sdd, cut by
tools/make_sdd_fixture.py) and on a synthetic multi-video tree."""
import os
import random

import numpy as np
import pytest

from desire_b200.utils.data_loader import DataLoader

HERE = os.path.dirname(os.path.abspath(__file__))
SDD = os.path.join(HERE, "golden", "sdd") + "/"


def literal_preprocess(csv_path, max_num_obj):
    """data_loader.py:98-144 restated: per frame, the objects in file order with their first (x, y)."""
    data = np.genfromtxt(csv_path, delimiter=",")
    frame_list = np.unique(data[0, :]).tolist()
    out = np.zeros((len(frame_list), max_num_obj, 3))
    for fi, frame in enumerate(frame_list):
        in_frame = data[:, data[0, :] == frame]
        rows = []
        for obj in in_frame[1, :].tolist():
            sel = in_frame[1, :] == obj
            rows.append([obj, in_frame[2, sel][0], in_frame[3, sel][0]])
        out[fi, :len(rows), :] = np.array(rows)
    return out, frame_list


def literal_window(current_data, idx, T, N):
    """data_loader.py:205-229 restated."""
    seq = current_data[idx:idx + T + 1]
    src_f, tgt_f = current_data[idx:idx + T], current_data[idx + 1:idx + T + 1]
    ids = np.unique(seq[:, :, 0])
    src, tgt = np.zeros((T, N, 3)), np.zeros((T, N, 3))
    for t in range(T):
        for k, oid in enumerate(ids):
            if oid == 0:
                continue
            s = src_f[t][src_f[t][:, 0] == oid]
            g = tgt_f[t][tgt_f[t][:, 0] == oid]
            if s.size:
                src[t, k] = s
            if g.size:
                tgt[t, k] = np.squeeze(g)
    return src, tgt


def test_preprocess_matches_literal(tmp_path):
    dl = DataLoader(2, 8, 40, 1, preprocess=True, data_dir=SDD, cache=False)
    ref, frames = literal_preprocess(os.path.join(SDD, "bookstore", "video0", "annotations_processed.csv"), 40)
    assert dl.frame_list[0] == frames
    assert np.array_equal(dl.data[0], ref)
    assert dl.num_obj_list[0][0] == 21
    assert dl.num_batches == 2 * int(int(40 / 10) / 2)


def test_next_batch_matches_literal_and_random_stream():
    dl = DataLoader(3, 8, 40, 1, preprocess=True, data_dir=SDD, cache=False)
    random.seed(5)
    xb, yb, dv = dl.next_batch()
    random.seed(5)
    idx = 0
    for x, y in zip(xb, yb):
        src, tgt = literal_window(dl.data[0], idx, 8, 40)
        assert np.array_equal(x, src) and np.array_equal(y, tgt)
        idx += random.randint(1, 8)
    assert dv == [0, 0, 0]
    # id 0 is a real SDD track id AND the "non-existent" sentinel (data_loader.py:221-222): its slot stays empty
    assert np.all(xb[0][:, 0, :] == 0)
    # target is the source shifted by one frame
    assert np.array_equal(xb[0][1:], yb[0][:-1])


def test_crowded_frame_raises_or_clips():
    with pytest.raises(ValueError):
        DataLoader(1, 8, 8, 1, preprocess=True, data_dir=SDD, cache=False)
    dl = DataLoader(1, 8, 8, 1, preprocess=True, data_dir=SDD, cache=False, clip=True)
    x, y, _ = dl.next_batch(random_update=False)
    assert x[0].shape == (8, 8, 3)
    assert (x[0][:, :, 0] != 0).sum() > 0


def test_pred_length_mode_and_model_layout():
    dl = DataLoader(2, 8, 40, 1, preprocess=True, data_dir=SDD, cache=False, pred_length=12)
    x, y, _ = dl.next_batch(random_update=False)
    assert x[0].shape == (8, 40, 3) and y[0].shape == (12, 40, 3)
    # the target continues the observed window: same row = same agent id
    ids_x = x[0][-1, :, 0]
    ids_y = y[0][0, :, 0]
    both = (ids_x != 0) & (ids_y != 0)
    assert both.sum() > 5 and np.array_equal(ids_x[both], ids_y[both])
    m = DataLoader.to_model_layout(x)
    assert m.shape == (2, 40, 8, 3) and m.dtype == np.float32
    assert np.array_equal(m[1, 3], x[1][:, 3].astype(np.float32))


def test_dataset_walk_wrap_and_leave_dataset(tmp_path):
    rng = np.random.default_rng(0)
    for v, nf in (("a/video0", 30), ("b/video0", 25)):
        d = tmp_path / v
        d.mkdir(parents=True)
        cols = []
        for oid in range(1, 5):
            for f in range(nf):
                cols.append((f, oid, rng.random() * 100, rng.random() * 100))
        arr = np.array(cols).T
        np.savetxt(d / "annotations_processed.csv", arr, delimiter=",")
    one = DataLoader(2, 8, 6, 1, preprocess=True, data_dir=str(tmp_path) + "/", cache=False)
    assert len(one.data) == 1                       # "leave_dataset" = take the first k datasets (data_loader.py:91)
    two = DataLoader(2, 8, 6, 2, preprocess=True, data_dir=str(tmp_path) + "/", cache=False)
    assert len(two.data) == 2
    seen = set()
    for _ in range(12):
        _, _, dv = two.next_batch(random_update=False)
        seen |= set(dv)
    assert seen == {0, 1}                           # pointer ticks to the next dataset and wraps (data_loader.py:249-258)


def test_scene_images_loaded_resized_and_blank_when_absent(tmp_path):
    """reference.jpg next to a video's CSV becomes the scene CNN input; datasets without one get zeros."""
    from PIL import Image
    from desire_b200.utils.data_loader import DataLoader
    rows = np.array([[0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9],
                     [1, 2] * 10, np.arange(20) * 1.0, np.arange(20) * 2.0])
    for name, with_img in (("a", True), ("b", False)):
        d = tmp_path / name / "video0"
        d.mkdir(parents=True)
        np.savetxt(d / "annotations_processed.csv", rows, delimiter=",")
        if with_img:
            img = np.zeros((10, 20, 3), np.uint8)
            img[:, :10, 0] = 255                                   # left half red
            Image.fromarray(img).save(d / "reference.png")
    dl = DataLoader(1, 2, 4, 2, preprocess=True, data_dir=str(tmp_path) + "/", cache=False, pred_length=3)
    files = dl._csv_files()
    idx_with = [i for i, f in enumerate(files) if "/a/" in f][0]
    sc = dl.scene_images([idx_with, 1 - idx_with], 8)
    assert sc.shape == (2, 8, 8, 3) and sc.dtype == np.float32
    assert sc[0, :, :3, 0].min() > 0.9 and sc[0, :, 5:, 0].max() < 0.1 and sc[0, ..., 1:].max() == 0
    assert sc[1].max() == 0


def test_seeded_loaders_walk_the_same_windows_and_state_restores_the_order():
    """Two ranks with the same seed must cut the same minibatches (train.py under torchrun), and a checkpointed
    loader state must replay the same order (--resume)."""
    a = DataLoader(2, 8, 40, 1, data_dir=SDD, pred_length=12, cache=False, seed=11)
    b = DataLoader(2, 8, 40, 1, data_dir=SDD, pred_length=12, cache=False, seed=11)
    for _ in range(3):
        xa, ya, da = a.next_batch()
        xb, yb, db = b.next_batch()
        assert all(np.array_equal(p, q) for p, q in zip(xa, xb)) and da == db
    st = a.state()
    nxt = a.next_batch()
    c = DataLoader(2, 8, 40, 1, data_dir=SDD, pred_length=12, cache=False, seed=999)
    c.set_state(st)
    again = c.next_batch()
    assert all(np.array_equal(p, q) for p, q in zip(nxt[0], again[0]))
    assert all(np.array_equal(p, q) for p, q in zip(nxt[1], again[1]))


def test_fix_id0_keeps_track_zero_and_auto_normalisation_maps_into_the_unit_square(tmp_path):
    root = tmp_path / "data"
    v = root / "scene" / "video0"
    v.mkdir(parents=True)
    # 30 frames, tracks 0 and 5 everywhere (SDD ids start at 0): frame, id, x, y rows as scripts/preprocess.py writes
    frames = np.repeat(np.arange(30), 2)
    ids = np.tile([0, 5], 30)
    xs = 100.0 + frames * 3 + ids
    ys = 700.0 - frames * 2 + ids
    np.savetxt(v / "annotations_processed.csv", np.stack([frames, ids, xs, ys]), delimiter=",")
    plain = DataLoader(1, 8, 4, 1, data_dir=str(root) + "/", pred_length=12, cache=False, seed=0)
    x, _, _ = plain.next_batch()
    assert set(np.unique(x[0][:, :, 0])) == {0.0, 5.0}
    assert (x[0][:, :, 0] == 5).sum() == 8                 # the reference quirk: track 0 is invisible (data_loader.py:221-222)
    assert np.count_nonzero(x[0][:, :, 1]) == 8
    fixed = DataLoader(1, 8, 4, 1, data_dir=str(root) + "/", pred_length=12, cache=False, seed=0, fix_id0=True,
                       normalize="auto")
    x, y, _ = fixed.next_batch()
    assert set(np.unique(x[0][:, :, 0])) == {0.0, 1.0, 6.0}      # ids shifted by one, 0 = empty slot only
    assert np.count_nonzero(x[0][:, :, 1]) == 16
    assert 0.0 < x[0][:, :, 1:][x[0][:, :, 0] > 0].min() and max(x[0][:, :, 1:].max(), y[0][:, :, 1:].max()) <= 1.0


def test_prefetch_thread_yields_the_same_batches_in_pinned_layout():
    a = DataLoader(2, 8, 40, 1, data_dir=SDD, pred_length=12, cache=False, seed=4)
    b = DataLoader(2, 8, 40, 1, data_dir=SDD, pred_length=12, cache=False, seed=4)
    got = []
    for x, y, sc, dval, st in a.prefetch_epoch(3, scene_size=16, depth=2):
        got.append((x.numpy().copy(), y.numpy().copy(), sc.numpy().copy(), list(dval), st))
    assert len(got) == 3
    for g in got:
        xb, yb, db = b.next_batch()
        assert np.array_equal(g[0], DataLoader.to_model_layout(xb)) and np.array_equal(g[1], DataLoader.to_model_layout(yb))
        assert g[0].shape == (2, 40, 8, 3) and g[1].shape == (2, 40, 12, 3) and g[2].shape == (2, 16, 16, 3)
        assert g[3] == db and g[4]["frame_pointer"] == b.frame_pointer
    # a rank's shard of the same minibatches: scenes rank, rank+world, ...
    c = DataLoader(2, 8, 40, 1, data_dir=SDD, pred_length=12, cache=False, seed=4)
    sh = [x.numpy().copy() for x, *_ in c.prefetch_epoch(3, depth=1, shard=(1, 2))]
    assert all(np.array_equal(s[0], g[0][1]) and s.shape[0] == 1 for s, g in zip(sh, got))
    # leaving the generator early stops the thread
    gen = a.prefetch_epoch(50, depth=1)
    next(gen)
    gen.close()
