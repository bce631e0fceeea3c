"""Gating-matrix files: the part of the reference's data module that is on the ELBO path.
Mirrors CelebAReader.set_gating_prob's load branch (utils_data.py:147-152) and the learned-mu
load of Learner.load_model (gated_ccvae.py:161-165).  Row index = z_c dim i, column = label j;
no transpose on load; float64 on disk, cast to fp32 by CCVAE.initialise_mu."""
from __future__ import annotations

import os

import numpy as np

CELEBA_EASY_LABELS = ['Arched_Eyebrows', 'Bags_Under_Eyes', 'Bangs', 'Black_Hair', 'Blond_Hair', 'Brown_Hair',
                      'Bushy_Eyebrows', 'Chubby', 'Eyeglasses', 'Heavy_Makeup', 'Male', 'No_Beard', 'Pale_Skin',
                      'Receding_Hairline', 'Smiling', 'Wavy_Hair', 'Wearing_Necktie', 'Young']


def gating_matrix_path(root, sup_frac):
    return os.path.join(root, "gating_matrix_{}.npy".format(sup_frac))


def load_gating_matrix(root, sup_frac):
    """-> float64 ndarray [18,18] exactly as stored (utils_data.py:149-152)."""
    path = gating_matrix_path(root, sup_frac)
    if not os.path.exists(path):
        raise FileNotFoundError("No gating matrix found at {} (generation from label co-occurrence "
                                "is outside the ELBO path)".format(path))
    mu = np.load(path)
    if mu.ndim != 2 or mu.shape[0] != mu.shape[1]:
        raise ValueError("gating matrix {} has shape {}".format(path, mu.shape))
    return mu


def load_learned_gating_matrix(param_dir, model_id):
    """gated_ccvae.py:161-163: learned mu saved by training ('best' | 'last')."""
    return np.load(os.path.join(param_dir, "learned_gating_matrix_{}.npy".format(model_id)))


def grouped_indices_from_labels(data):
    """rows of a binary label matrix [N, n_labels] -> list of the positive label indices of every row that has at
    least one (utils_data.py:163-165: np.nonzero + cut at row changes; rows without a positive label form no group)."""
    data = np.asarray(data)
    where_one_x, where_one_y = np.nonzero(data)
    cut_idx = np.flatnonzero(np.r_[True, where_one_x[1:] != where_one_x[:-1], True])
    return [where_one_y[i:j] for i, j in zip(cut_idx[:-1], cut_idx[1:])]


def cooccurrence_counts(data):
    """C[i, j] = number of rows with labels i and j both positive, 0 on the diagonal; n = rows with any positive label.
    The double loop of utils.py:137-144 as one integer matrix product (exact)."""
    y = (np.asarray(data) != 0).astype(np.int64)
    counts = y.T @ y
    np.fill_diagonal(counts, 0)
    return counts, int((y.sum(axis=1) > 0).sum())


def create_gating_matrix(grouped_indices, n_labels):
    """utils.py:132-149: co-occurrence counts of the label groups / number of groups, diagonal set to 1 (float64)."""
    n_elems = len(grouped_indices)
    y = np.zeros((n_elems, n_labels), dtype=np.int64)
    for r, group in enumerate(grouped_indices):
        y[r, np.asarray(group, dtype=np.int64)] = 1
    counts = y.T @ y
    np.fill_diagonal(counts, 0)
    gating_matrix = counts.astype(np.float64) / n_elems
    np.fill_diagonal(gating_matrix, 1)
    return gating_matrix


def initial_gating_matrix(sup_frac, sup_labels=None, valid_labels=None, n_labels=len(CELEBA_EASY_LABELS)):
    """the generate branch of CelebAReader.set_gating_prob (utils_data.py:153-168): 0.5 off the diagonal when there
    is no supervision, else the co-occurrence statistics of the supervised + validation label rows."""
    if sup_frac == 0.0:
        mu = np.ones((n_labels, n_labels)) / 2.0
        np.fill_diagonal(mu, 1.)
        return mu
    data = np.concatenate((np.asarray(sup_labels), np.asarray(valid_labels)), axis=0)
    return create_gating_matrix(grouped_indices_from_labels(data), n_labels)


def gating_matrix_csv(mu, columns=CELEBA_EASY_LABELS):
    """text of `pd.DataFrame(mu, index=z1.., columns=labels).to_csv()` (utils_data.py:172-174, gated_ccvae.py:399-403):
    shortest round-trip repr of every entry in the array's own precision."""
    mu = np.asarray(mu)
    lines = ["," + ",".join(columns)]
    for i in range(mu.shape[0]):
        lines.append("z{},".format(i + 1) + ",".join(str(t) for t in mu[i]))
    return "\n".join(lines) + "\n"


class GatingMatrixReader:
    """The slice of CelebAReader the Learner consumes: `.init_gating_prob` (gated_ccvae.py:506), loaded from
    `root/gating_matrix_{sup_frac}.npy` if it exists, else generated from label rows and saved as `.npy` + `.csv`
    (utils_data.py:147-176)."""

    def __init__(self, root, sup_frac, batch_size=None, sup_labels=None, valid_labels=None):
        self.root, self.sup_frac, self.batch_size = root, sup_frac, batch_size
        self.set_gating_prob(sup_labels, valid_labels)

    def set_gating_prob(self, sup_labels=None, valid_labels=None):
        path = gating_matrix_path(self.root, self.sup_frac)
        if os.path.exists(path):
            self.init_gating_prob = load_gating_matrix(self.root, self.sup_frac)
            return
        if self.sup_frac != 0.0 and (sup_labels is None or valid_labels is None):
            raise FileNotFoundError("No gating matrix found at {} and no label rows to generate it from".format(path))
        mu = initial_gating_matrix(self.sup_frac, sup_labels, valid_labels)
        self.init_gating_prob = mu
        os.makedirs(self.root, exist_ok=True)
        np.save(path, mu)
        with open(os.path.join(self.root, "gating_matrix_{}.csv".format(self.sup_frac)), "w") as fh:
            fh.write(gating_matrix_csv(mu))


class SyntheticReader:
    """Stand-in for CelebAReader (utils_data.py:31-176) with the interface the Learner consumes: `.n_s` samples,
    `.step()` yielding `(xs, ys)` batches forever (xs uint8 or fp32 [B,64,64,3], ys int64 [B,18] or None).  CelebA
    itself is not available; images are seeded synthetic 8-bit pixels, labels Bernoulli(0.5)."""

    def __init__(self, n_samples, batch_size, supervised=True, seed=0, dtype="uint8", y_dim=18):
        import torch
        self.n_s, self.batch_size, self.supervised = int(n_samples), int(batch_size), bool(supervised)
        gen = torch.Generator().manual_seed(seed)
        self.x = torch.randint(0, 256, (self.n_s, 64, 64, 3), generator=gen, dtype=torch.uint8)
        if dtype != "uint8":
            self.x = self.x.to(torch.float32) / 255.0
        self.y = (torch.rand(self.n_s, y_dim, generator=gen) < 0.5).to(torch.int64)

    def step(self):
        i = 0
        while True:
            idx = [(i + j) % self.n_s for j in range(self.batch_size)]
            i = (i + self.batch_size) % self.n_s
            yield self.x[idx], (self.y[idx] if self.supervised else None)


# ---------------------------------------------------------------------------------------------------------------------
# CelebA input pipeline (utils_data.py:31-198).  Host side of the path: it produces what `Learner.train_step` consumes.
# The image blobs are not part of the reference checkout, so this is exercised on a synthetic directory of the same
# layout (tests/test_celeba_reader.py).  One addition: `dtype="uint8"` hands the resized 8-bit pixels through unchanged -
# the tensor-core engine normalises u8 / 255 on the device, bit-exactly as `np.float32(img) / 255.0` (:57-59) does, and
# the batch that crosses PCIe is 4x smaller.
# ---------------------------------------------------------------------------------------------------------------------
CELEBA_LABELS = ['5_o_Clock_Shadow', 'Arched_Eyebrows', 'Attractive', 'Bags_Under_Eyes', 'Bald', 'Bangs', 'Big_Lips',
                 'Big_Nose', 'Black_Hair', 'Blond_Hair', 'Blurry', 'Brown_Hair', 'Bushy_Eyebrows', 'Chubby', 'Double_Chin',
                 'Eyeglasses', 'Goatee', 'Gray_Hair', 'Heavy_Makeup', 'High_Cheekbones', 'Male', 'Mouth_Slightly_Open',
                 'Mustache', 'Narrow_Eyes', 'No_Beard', 'Oval_Face', 'Pale_Skin', 'Pointy_Nose', 'Receding_Hairline',
                 'Rosy_Cheeks', 'Sideburns', 'Smiling', 'Straight_Hair', 'Wavy_Hair', 'Wearing_Earrings', 'Wearing_Hat',
                 'Wearing_Lipstick', 'Wearing_Necklace', 'Wearing_Necktie', 'Young']
CELEBA_SPLIT = {"train": 162770, "valid": 19867, "test": 19962}


class LabelTable:
    """image file names + their binary label rows (the reference's CSV namedtuple: header, index, data)."""

    def __init__(self, header, index, data):
        self.header, self.index, self.data = list(header), list(index), np.asarray(data)

    def rows(self, lo, hi=None):
        return LabelTable(self.header, self.index[lo:hi], self.data[lo:hi])

    def __len__(self):
        return len(self.index)


class DataLoader:
    """utils_data.py:31-88: endless batches of (images [B,64,64,3], labels [B,18]) in a (shuffled) fixed order that wraps
    around; `.n_s` samples; `.step()` yields the batch read last and reads the next one; `.reset()` rewinds."""

    def __init__(self, data_dir, cached_data, batch_size, shuffle=True, dtype="float32", rng=None):
        self.data_dir, self.cached_data, self.bs = data_dir, cached_data, int(batch_size)
        self.n_s = len(cached_data.data)
        self.dtype = dtype
        self.idxs = list(range(self.n_s))
        if shuffle:
            (rng if rng is not None else np.random).shuffle(self.idxs)
        self.start = 0
        self.Xs, self.ys = self.read_data(self.get_batch())

    def read_data(self, idxs, normalise=True):
        from PIL import Image
        out = np.empty((len(idxs), 64, 64, 3), dtype=np.uint8)
        for r, i in enumerate(idxs):
            with Image.open(os.path.join(self.data_dir, self.cached_data.index[i])) as img:
                out[r] = np.asarray(img.convert("RGB").resize((64, 64)))      # PIL's default resampling, as :54-56
        labels = self.cached_data.data[idxs]
        if self.dtype == "uint8":
            return out, labels
        X = out.astype(np.float32)
        return (X / 255.0 if normalise else X), labels

    def get_batch(self):
        n, s = self.n_s, self.start
        if s + self.bs < n:
            batch = self.idxs[s:s + self.bs]
            self.start = s + self.bs
        else:           # wrap around: the tail of the order followed by its head
            batch = self.idxs[s:] + self.idxs[:self.bs - (n - s)]
            self.start = (s + self.bs) % n
        return batch

    def step(self):
        while True:
            yield self.Xs, self.ys
            self.Xs, self.ys = self.read_data(self.get_batch())

    def reset(self):
        self.start = 0


class CelebAReader(GatingMatrixReader):
    """utils_data.py:91-198: `list_attr_celeba.csv` (40 attributes, -1 / 1) -> the 18 easy labels (0 / 1); train / valid /
    test by position; the first `sup_frac` of the training rows are the supervised set; loaders over
    `root/img_align_celeba`; `init_gating_prob` loaded or generated from the supervised + validation label rows."""

    def __init__(self, root, sup_frac, batch_size, split_map=None, dtype="float32"):
        self.root, self.sup_frac, self.batch_size, self.dtype = root, sup_frac, batch_size, dtype
        self.split_map = dict(split_map or CELEBA_SPLIT)
        self.sub_label_inds = [i for i, name in enumerate(CELEBA_LABELS) if name in CELEBA_EASY_LABELS]
        self.attr = self._load_csv("list_attr_celeba.csv", header=0)
        self.init_gating_prob = None

    def _load_csv(self, filename, header=None):
        import csv
        with open(os.path.join(self.root, filename)) as fh:
            rows = [r for r in csv.reader(fh) if r]
        if header is not None:
            rows = rows[header + 1:]
        index = [r[0] for r in rows]
        data = np.array([[int(v) for v in r[1:]] for r in rows], dtype=np.int64).reshape(len(rows), -1)
        data = (data == 1).astype(np.int64)[:, self.sub_label_inds]           # -1 -> 0, 1 -> 1, easy labels only
        return LabelTable(["image_id"] + CELEBA_EASY_LABELS, index, data)

    def load_split_data(self):
        n_train, n_valid = self.split_map["train"], self.split_map["valid"]
        train = self.attr.rows(0, n_train)
        cached = {"train": train}
        if self.sup_frac == 0.0:
            cached["unsup"] = train
        elif self.sup_frac == 1.0:
            cached["sup"] = train
        else:
            n_sup = int(n_train * self.sup_frac)
            cached["sup"], cached["unsup"] = train.rows(0, n_sup), train.rows(n_sup)
        cached["valid"] = self.attr.rows(n_train, n_train + n_valid)
        cached["test"] = self.attr.rows(n_train + n_valid)
        return cached

    def setup_data_loaders(self, shuffle=True, rng=None):
        if self.sup_frac == 0.0:
            modes = ["unsup", "test"]
        elif self.sup_frac == 1.0:
            modes = ["sup", "test", "valid"]
        else:
            modes = ["unsup", "test", "sup", "valid"]
        cached = self.load_split_data()
        self.set_gating_prob(cached["sup"].data if "sup" in cached else None, cached["valid"].data)
        img_dir = os.path.join(self.root, "img_align_celeba")
        return {m: DataLoader(img_dir, cached[m], self.batch_size, shuffle=shuffle, dtype=self.dtype, rng=rng) for m in modes}


class PrefetchLoader:
    """Same batches, same order as the `DataLoader` it wraps, but decoded ahead of the consumer: the reference reads and
    resizes every image synchronously between two train steps (utils_data.py:48-63), which bounds its epoch at ~700
    images/s - three orders of magnitude below the step.  Here `depth` batches are in flight, each
    image decoded by a pool of `workers` threads (PIL releases the GIL while decoding and resizing), and the batches
    come out as uint8 (optionally pinned) arrays ready for `Learner.train_step`."""

    def __init__(self, loader, workers=8, depth=4, pin=False):
        import concurrent.futures as cf
        self.loader, self.n_s, self.depth, self.pin = loader, loader.n_s, max(1, int(depth)), bool(pin)
        self.pool = cf.ThreadPoolExecutor(max_workers=max(1, int(workers)))
        from collections import deque
        self.queue = deque()             # decode jobs submitted ahead, in the wrapped loader's order
        self.Xs = self.ys = None         # the batch handed out last (DataLoader.Xs / .ys)

    def _decode_one(self, name):
        from PIL import Image
        with Image.open(os.path.join(self.loader.data_dir, name)) as img:
            return np.asarray(img.convert("RGB").resize((64, 64)))

    def _submit(self, idxs):
        names = [self.loader.cached_data.index[i] for i in idxs]
        return [self.pool.submit(self._decode_one, n) for n in names], self.loader.cached_data.data[idxs]

    def _finish(self, job):
        futures, labels = job
        X = np.stack([f.result() for f in futures]).astype(np.uint8, copy=False)
        if self.loader.dtype != "uint8":
            X = X.astype(np.float32) / 255.0
        if self.pin:
            import torch
            X = torch.from_numpy(X).pin_memory()
        return X, labels

    def step(self):
        """the wrapped loader's sequence, with the wrapped loader's `step()` semantics (utils_data.py:77-80): every call
        first yields the batch read LAST (the loader's constructor batch on the first call, the batch the previous
        `step()` iterator handed out last afterwards) and then continues with `get_batch()` after `get_batch()`.  The
        read-ahead queue and the last batch live on this object, so batches submitted ahead by one iterator are served
        by the next one instead of being dropped."""
        if self.Xs is None:
            first = (self.loader.Xs, self.loader.ys)
            if self.pin:
                import torch
                first = (torch.from_numpy(np.ascontiguousarray(first[0])).pin_memory(), first[1])
            self.Xs, self.ys = first
        while True:
            yield self.Xs, self.ys
            while len(self.queue) < self.depth:
                self.queue.append(self._submit(self.loader.get_batch()))
            job = self.queue.popleft()
            self.queue.append(self._submit(self.loader.get_batch()))
            self.Xs, self.ys = self._finish(job)

    def reset(self):
        """utils_data.py:86-87 rewinds the read position only; batches already submitted ahead are discarded so that the
        next read is the first batch of the order, as it is for the wrapped loader."""
        for futures, _ in self.queue:
            for f in futures:
                f.cancel()
        self.queue.clear()
        self.loader.reset()

    def close(self):
        self.pool.shutdown(wait=False, cancel_futures=True)


class CachedLoader:
    """Decode once, serve from memory.  JPEG decode + resize runs at ~700-1000 images/s per host however it is threaded
    (measured), the ELBO step at 1.4 M images/s per GPU - so the only loader that keeps up is one that does
    not decode: all `n_s` images of the wrapped `DataLoader`'s table are decoded ONCE to a uint8 array [n_s,64,64,3]
    (12 KB per image: the 162 770 CelebA training images are 2.0 GB), optionally kept as a `.npy` next to the data and
    optionally resident on the GPU (`device=`: 2 GB of 180 GB, batches are then gathered on the device and never cross
    PCIe).  Batches follow the wrapped loader's own sequence (same shuffled order, same wrap-around) and equal its
    batches bit for bit."""

    def __init__(self, loader, cache_path=None, workers=8, device=None):
        self.loader, self.n_s = loader, loader.n_s
        if cache_path is not None and os.path.exists(cache_path):
            pixels = np.load(cache_path, mmap_mode=None)
            if pixels.shape != (self.n_s, 64, 64, 3) or pixels.dtype != np.uint8:
                raise ValueError("{}: cached pixels {} {} do not match the table ({} images)".format(
                    cache_path, pixels.shape, pixels.dtype, self.n_s))
        else:
            pre = PrefetchLoader(loader, workers=workers, depth=1)
            jobs = [pre._submit(list(range(i, min(i + 256, self.n_s)))) for i in range(0, self.n_s, 256)]
            pixels = np.concatenate([np.stack([f.result() for f in futs]) for futs, _ in jobs]).astype(np.uint8, copy=False)
            pre.close()
            if cache_path is not None:
                tmp = cache_path + ".tmp.npy"
                np.save(tmp, pixels)
                os.replace(tmp, cache_path)
        self.labels = np.asarray(loader.cached_data.data)
        self.device = device
        if device is not None:
            import torch
            self.pixels = torch.from_numpy(pixels).to(device)
        else:
            self.pixels = pixels
        self.Xs = self.ys = None         # the batch handed out last (DataLoader.Xs / .ys)

    def _gather(self, idxs):
        if self.device is not None:
            import torch
            x = self.pixels.index_select(0, torch.as_tensor(idxs, dtype=torch.long, device=self.pixels.device))
        else:
            x = self.pixels[idxs]
            if self.loader.dtype != "uint8":
                x = x.astype(np.float32) / 255.0
        return x, self.labels[idxs]

    def step(self):
        """`DataLoader.step()` semantics (utils_data.py:77-80): yield the batch read last, then read on.  The first
        batch is the one the wrapped loader drew in its constructor: indices order[0 : bs] (its start was 0)."""
        if self.Xs is None:
            order, bs = self.loader.idxs, self.loader.bs
            first = order[:bs] if bs < self.n_s else order[:] + order[:bs - self.n_s]
            self.Xs, self.ys = self._gather(first)
        while True:
            yield self.Xs, self.ys
            self.Xs, self.ys = self._gather(self.loader.get_batch())

    def reset(self):
        self.loader.reset()
