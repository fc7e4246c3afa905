"""DataLoader — same constructor, attributes and methods as the reference's utils/data_loader.py:20-266
(`DataLoader(batch_size, seq_length, max_num_obj, leave_dataset, preprocess)`, `next_batch`,
`tick_batch_pointer`, `reset_batch_pointer`, `num_batches`), vectorised.

What is kept exactly (tests/test_data_loader.py checks it against a literal restatement):
  * CSV format of scripts/preprocess.py:30-34 — 4 rows: frame ids, object ids, x, y (raw pixels);
  * datasets are taken in os.walk order and only the first `leave_dataset` of them (data_loader.py:88-92 —
    the flag is NOT leave-one-out in the reference);
  * per-dataset array [frames, max_num_obj, 3] = (id, x, y), objects in file order (:113-141);
  * next_batch: window of seq_length+1 frames, source = first T, target = shifted by ONE frame (:205-207),
    output row = index of the id in the sorted unique id list of the window INCLUDING id 0 (:209-229),
    id 0 never written (:221-222), frame pointer += random.randint(1, T) (:236), dataset wrap (:249-258),
    num_batches = 2 * int(sum(len/(T+2)) / batch) (:171-183).
What is added: `pred_length` — when set, the target holds the NEXT pred_length frames after the
observed window (what the model's future encoder and losses need); `data_dir`; `clip` — the reference raises
when a frame holds more than max_num_obj objects (10 of the 60 SDD videos do), clip=True keeps the
first max_num_obj instead; `normalize=(W,H)` divides pixel coordinates ("auto" = the data's own extent, so that
positions land in the unit square the IOC stage's scene gather and log-polar radii assume); `seed` — a private
random.Random for the frame-pointer stride (the reference draws from Python's global, unseeded `random`, :236), so
every rank of a data-parallel job walks the same windows and a checkpoint can restore the order (state()/set_state());
`fix_id0` — SDD track ids start at 0 and 0 is also the "no object" sentinel (:221-222, model/model.py:206,357), so
the reference silently drops track 0 of every video; fix_id0=True shifts the real ids by +1 when the cache is loaded;
`prefetch_epoch()` — a background thread builds batch t+1 (windows, agent-major layout, scene images) into pinned
host buffers while the GPU runs step t.

The reference's frame_preprocess is O(frames x annotations) boolean masking (4.8e9 compares for
bookstore/video0); here one stable sort by frame + a rank-within-frame scatter does it.
"""
from __future__ import annotations

import os
import pickle
import queue
import random
import threading

import numpy as np


class DataLoader(object):
    def __init__(self, batch_size=50, seq_length=5, max_num_obj=40, leave_dataset=1, preprocess=False,
                 data_dir="data/", pred_length=None, clip=False, normalize=None, cache=True, seed=None,
                 fix_id0=False):
        self.leave_dataset = leave_dataset
        self.data_dir = data_dir
        self.frame_pointer = 0
        self.dataset_pointer = 0
        self.max_num_obj = max_num_obj
        self.batch_size = batch_size
        self.seq_length = seq_length
        self.pred_length = pred_length
        self.clip = clip
        self.normalize = normalize
        self.fix_id0 = bool(fix_id0)
        # seed=None keeps the reference's behaviour (the process-global `random`, data_loader.py:236)
        self.rng = random.Random(seed) if seed is not None else random
        data_file = os.path.join(self.data_dir, "trajectories.cpkl")
        # the reference ALWAYS re-runs the preprocessing (its guard is commented out, :53-59); here the cache is
        # reused unless preprocess=True or it is missing/stale
        if preprocess or not cache or not self._cache_ok(data_file):
            self.frame_preprocess(data_file)
        self.load_preprocessed(data_file)
        self.reset_batch_pointer()

    # ------------------------------------------------------------------ preprocessing
    def _csv_files(self):
        out = []
        for subdir, _dirs, files in os.walk(self.data_dir):
            for f in files:
                if f == "annotations_processed.csv":
                    out.append(os.path.join(subdir, f))
        return out[: self.leave_dataset]

    def _cache_ok(self, data_file):
        if not os.path.exists(data_file):
            return False
        try:
            with open(data_file, "rb") as fh:
                raw = pickle.load(fh)
            return len(raw) == 4 and raw[3] == (self._csv_files(), self.max_num_obj, self.clip)
        except Exception:
            return False

    def frame_preprocess(self, data_file):
        all_frame_data, frame_list_data, num_obj_data = [], [], []
        for path in self._csv_files():
            data = np.loadtxt(path, delimiter=",", ndmin=2)
            frames, ids, xs, ys = data[0], data[1], data[2], data[3]
            order = np.argsort(frames, kind="stable")                 # file order kept inside a frame
            f_sorted = frames[order]
            frame_list, start, counts = np.unique(f_sorted, return_index=True, return_counts=True)
            rank = np.arange(f_sorted.size) - np.repeat(start, counts)  # position of the annotation in its frame
            fidx = np.repeat(np.arange(frame_list.size), counts)
            if counts.max() > self.max_num_obj and not self.clip:
                raise ValueError("%s: a frame holds %d objects > max_num_obj=%d (pass clip=True to keep the first %d)"
                                 % (path, counts.max(), self.max_num_obj, self.max_num_obj))
            keep = rank < self.max_num_obj
            arr = np.zeros((frame_list.size, self.max_num_obj, 3))
            arr[fidx[keep], rank[keep], 0] = ids[order][keep]
            arr[fidx[keep], rank[keep], 1] = xs[order][keep]
            arr[fidx[keep], rank[keep], 2] = ys[order][keep]
            all_frame_data.append(arr)
            frame_list_data.append(frame_list.tolist())
            num_obj_data.append(counts.tolist())
        os.makedirs(os.path.dirname(os.path.abspath(data_file)), exist_ok=True)
        tmp = "%s.tmp.%d" % (data_file, os.getpid())          # a concurrent reader never sees a partial file
        with open(tmp, "wb") as fh:
            pickle.dump((all_frame_data, frame_list_data, num_obj_data,
                         (self._csv_files(), self.max_num_obj, self.clip)), fh, protocol=2)
        os.replace(tmp, data_file)

    def load_preprocessed(self, data_file):
        with open(data_file, "rb") as fh:
            self.raw_data = pickle.load(fh)
        self.data = self.raw_data[0]
        self.frame_list = self.raw_data[1]
        self.num_obj_list = self.raw_data[2]
        if self.fix_id0:
            # slots [0, count_f) of frame f hold real objects (frame_preprocess fills them in file order): shift their
            # ids so that a real track 0 is no longer mistaken for an empty slot
            fixed = []
            for d, counts in zip(self.data, self.num_obj_list):
                d = d.copy()
                real = np.arange(d.shape[1])[None, :] < np.minimum(np.asarray(counts), d.shape[1])[:, None]
                d[..., 0] = np.where(real, d[..., 0] + 1, d[..., 0])
                fixed.append(d)
            self.data = fixed
        if isinstance(self.normalize, str):
            if self.normalize != "auto":
                raise ValueError("normalize must be None, (W, H) or 'auto'")
            w = max([float(d[..., 1].max()) for d in self.data] + [1.0])
            h = max([float(d[..., 2].max()) for d in self.data] + [1.0])
            self.normalize = (float(np.ceil(w)), float(np.ceil(h)))
        if self.normalize is not None:
            w, h = self.normalize
            self.data = [np.concatenate([d[..., :1], d[..., 1:2] / w, d[..., 2:3] / h], -1) for d in self.data]
        counter = 0
        for all_frame_data in self.data:
            counter += int(len(all_frame_data) / (self.seq_length + 2))
        self.num_batches = int(counter / self.batch_size) * 2

    # ------------------------------------------------------------------ batching
    def _window(self, current_data, idx):
        """-> (source [T,N,3], target [T or Tf,N,3]) for the window starting at frame idx."""
        T, N = self.seq_length, self.max_num_obj
        n_tgt = T if self.pred_length is None else self.pred_length
        src_frames = current_data[idx:idx + T]
        if self.pred_length is None:
            seq = current_data[idx:idx + T + 1]
            tgt_frames = current_data[idx + 1:idx + T + 1]
        else:
            seq = current_data[idx:idx + T + self.pred_length]
            tgt_frames = current_data[idx + T:idx + T + self.pred_length]
        uniq = np.unique(seq[:, :, 0])
        if uniq.shape[0] > N:
            if not self.clip:
                raise ValueError("window at frame %d has %d unique ids > max_num_obj=%d" % (idx, uniq.shape[0], N))
            uniq = uniq[:N]

        def place(frames, n):
            out = np.zeros((n, N, 3))
            ids = frames[:, :, 0]
            row = np.searchsorted(uniq, ids)
            ok = (ids != 0) & (row < uniq.shape[0])
            ok &= uniq[np.minimum(row, uniq.shape[0] - 1)] == ids
            t_idx = np.broadcast_to(np.arange(n)[:, None], ids.shape)
            out[t_idx[ok], row[ok]] = frames[ok]
            return out

        return place(src_frames, T), place(tgt_frames, n_tgt)

    def next_batch(self, random_update=True):
        x_batch, y_batch, dval = [], [], []
        need = self.seq_length if self.pred_length is None else self.seq_length + self.pred_length - 1
        i = 0
        guard = 0
        while i < self.batch_size:
            current_data = self.data[self.dataset_pointer]
            idx = self.frame_pointer
            if idx + need < current_data.shape[0]:
                src, tgt = self._window(current_data, idx)
                x_batch.append(src)
                y_batch.append(tgt)
                if random_update:
                    self.frame_pointer += self.rng.randint(1, self.seq_length)
                else:
                    self.frame_pointer += self.seq_length
                dval.append(self.dataset_pointer)
                i += 1
                guard = 0
            else:
                self.tick_batch_pointer()
                guard += 1
                if guard > len(self.data):
                    raise ValueError("no dataset holds a window of %d frames" % (need + 1))
        return x_batch, y_batch, dval

    def tick_batch_pointer(self):
        self.dataset_pointer += 1
        self.frame_pointer = 0
        if self.dataset_pointer >= len(self.data):
            self.dataset_pointer = 0

    def reset_batch_pointer(self):
        self.dataset_pointer = 0
        self.frame_pointer = 0

    # ------------------------------------------------------------------ resumable order (checkpoints)
    def state(self):
        """Pointers + RNG state: what a checkpoint needs to replay the same sequence of minibatches."""
        rng = self.rng.getstate() if self.rng is not random else None
        return {"dataset_pointer": self.dataset_pointer, "frame_pointer": self.frame_pointer, "rng": rng}

    def set_state(self, st):
        self.dataset_pointer, self.frame_pointer = int(st["dataset_pointer"]), int(st["frame_pointer"])
        if st.get("rng") is not None and self.rng is not random:
            self.rng.setstate(st["rng"])

    # ------------------------------------------------------------------ background prefetch into pinned buffers
    def prefetch_epoch(self, n_batches, scene_size=None, depth=2, shard=None, random_update=True):
        """Generator over the next `n_batches` minibatches, built one ahead by a background thread:
            for x, y, scene, dval, state in loader.prefetch_epoch(nb, scene_size): ...
        x [B,N,Tp,3], y [B,N,Tf,3] (agent-major, float32) and scene [B,S,S,3] (None without scene_size) are torch
        tensors in pinned host memory (plain host memory when CUDA is absent) taken from a ring of depth+1 buffer
        sets, so the batch handed out stays valid until the next iteration asks for another one.  `shard` =
        (rank, world) keeps only this rank's scenes (round-robin, dist.shard_scenes); every rank must build its loader
        with the same seed so that all ranks cut the same minibatch.  `state` is loader.state() AFTER that batch —
        store it in a checkpoint to resume the order.  Only the background thread touches the pointers meanwhile."""
        import torch
        pin = torch.cuda.is_available()
        ring, q = {}, queue.Queue(maxsize=max(1, depth))
        stop = threading.Event()

        def buf(slot, name, shape):
            key = (slot, name)
            if key not in ring or tuple(ring[key].shape) != tuple(shape):
                ring[key] = torch.empty(shape, dtype=torch.float32, pin_memory=pin)
            return ring[key]

        def work():
            try:
                for i in range(n_batches):
                    if stop.is_set():
                        return
                    xval, yval, dval = self.next_batch(random_update)
                    x, y = self.to_model_layout(xval), self.to_model_layout(yval)
                    sc = self.scene_images(dval, scene_size) if scene_size else None
                    if shard is not None:
                        mine = list(range(shard[0], x.shape[0], shard[1]))
                        x, y = x[mine], y[mine]
                        sc = sc[mine] if sc is not None else None
                        dval = [dval[j] for j in mine]
                    slot = i % (depth + 1)
                    out = []
                    for name, a in (("x", x), ("y", y), ("scene", sc)):
                        if a is None:
                            out.append(None)
                            continue
                        t = buf(slot, name, a.shape)
                        t.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))
                        out.append(t)
                    q.put((out[0], out[1], out[2], dval, self.state()))
                q.put(None)
            except BaseException as e:      # surfaced in the consumer
                q.put(e)

        th = threading.Thread(target=work, name="desire-prefetch", daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while th.is_alive():            # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                th.join(timeout=0.05)

    # ------------------------------------------------------------------ scene context
    SCENE_IMAGE_NAMES = ("reference.jpg", "reference.png", "reference.jpeg")

    def scene_images(self, dval, size):
        """Scene images for the sequences of a batch: for every dataset index in `dval` (third value of next_batch)
        the `reference.jpg` that the Stanford Drone Dataset ships next to each video's annotations, resized to
        size x size, RGB float32 in [0,1] -> [B,size,size,3] (the scene CNN's input layout).  The reference repository
        ships only the CSVs (its data/ holds no images), so a missing file yields a blank (zero) scene — the same
        input the training loop used before.  Decoded images are cached per dataset."""
        if not hasattr(self, "_scene_cache"):
            self._scene_cache = {}
        files = self._csv_files()
        out = np.zeros((len(dval), size, size, 3), np.float32)
        for i, d in enumerate(dval):
            key = (int(d), int(size))
            if key not in self._scene_cache:
                img = None
                folder = os.path.dirname(files[int(d)]) if int(d) < len(files) else None
                if folder:
                    for name in self.SCENE_IMAGE_NAMES:
                        path = os.path.join(folder, name)
                        if os.path.exists(path):
                            from PIL import Image      # only needed when an image is actually present
                            with Image.open(path) as im:
                                im = im.convert("RGB").resize((size, size), Image.BILINEAR)
                                img = np.asarray(im, np.float32) / 255.0
                            break
                self._scene_cache[key] = img
            if self._scene_cache[key] is not None:
                out[i] = self._scene_cache[key]
        return out

    # ------------------------------------------------------------------ model-side layout
    @staticmethod
    def to_model_layout(batch):
        """list[B] of [T,N,3] (time-major, what next_batch returns) -> [B,N,T,3] float32 (agent-major, the
        layout of the model's placeholders, model/model.py:91-105; train.py:158-173 forgot this transpose)."""
        return np.ascontiguousarray(np.stack(batch).transpose(0, 2, 1, 3), dtype=np.float32)
