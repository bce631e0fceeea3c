"""CelebA input pipeline (utils_data.py:31-198) on a synthetic directory of the dataset's layout: label mapping, splits,
supervised fraction, batch order with wrap-around, resize, uint8 / float32 outputs, gating matrix generated and saved."""
import os

import numpy as np
import pytest

import gccvae_b200 as G
from gccvae_b200 import utils_data as UD

N, SPLIT = 40, {"train": 24, "valid": 8, "test": 8}


@pytest.fixture(scope="module")
def root(tmp_path_factory):
    from PIL import Image
    d = tmp_path_factory.mktemp("celeba")
    os.makedirs(d / "img_align_celeba")
    rng = np.random.default_rng(0)
    attrs = np.where(rng.random((N, 40)) < 0.35, 1, -1)
    with open(d / "list_attr_celeba.csv", "w") as fh:
        fh.write("image_id," + ",".join(UD.CELEBA_LABELS) + "\n")
        for i in range(N):
            name = "%06d.png" % (i + 1)
            fh.write(name + "," + ",".join(str(v) for v in attrs[i]) + "\n")
            Image.fromarray(rng.integers(0, 256, (109, 89, 3), dtype=np.uint8)).save(d / "img_align_celeba" / name)
    return str(d), attrs


def test_label_mapping_and_splits(root):
    path, attrs = root
    rdr = G.utils_data.CelebAReader(path, 0.25, 4, split_map=SPLIT)
    easy = [UD.CELEBA_LABELS.index(n) for n in UD.CELEBA_EASY_LABELS]
    assert rdr.sub_label_inds == easy and len(easy) == 18
    assert np.array_equal(rdr.attr.data, (attrs[:, easy] == 1).astype(np.int64))
    c = rdr.load_split_data()
    assert (len(c["train"]), len(c["sup"]), len(c["unsup"]), len(c["valid"]), len(c["test"])) == (24, 6, 18, 8, 8)
    assert c["sup"].index == ["%06d.png" % i for i in range(1, 7)] and c["test"].index[0] == "000033.png"
    assert set(G.utils_data.CelebAReader(path, 0.0, 4, split_map=SPLIT).load_split_data()) == {"train", "unsup", "valid", "test"}
    assert len(G.utils_data.CelebAReader(path, 1.0, 4, split_map=SPLIT).load_split_data()["sup"]) == 24


def test_loaders_batches_and_gating_matrix(root):
    from PIL import Image
    path, attrs = root
    for f in os.listdir(path):
        if f.startswith("gating_matrix"):
            os.remove(os.path.join(path, f))
    rdr = G.utils_data.CelebAReader(path, 0.25, 4, split_map=SPLIT, dtype="uint8")
    loaders = rdr.setup_data_loaders(shuffle=False)
    assert set(loaders) == {"unsup", "test", "sup", "valid"} and loaders["sup"].n_s == 6
    # gating matrix from the supervised + validation rows, saved next to the data
    c = rdr.load_split_data()
    want = UD.create_gating_matrix(UD.grouped_indices_from_labels(np.concatenate((c["sup"].data, c["valid"].data))), 18)
    assert rdr.init_gating_prob.tobytes() == want.tobytes()
    assert os.path.exists(os.path.join(path, "gating_matrix_0.25.npy")) and os.path.exists(os.path.join(path, "gating_matrix_0.25.csv"))
    # batch order of 6 samples in batches of 4, unshuffled: the reference's wrap-around rule
    it = iter(loaders["sup"].step())
    seen = [next(it) for _ in range(4)]
    idx = [[0, 1, 2, 3], [4, 5, 0, 1], [2, 3, 4, 5], [0, 1, 2, 3]]
    for (X, y), want_idx in zip(seen, idx):
        assert X.dtype == np.uint8 and X.shape == (4, 64, 64, 3) and y.shape == (4, 18)
        assert np.array_equal(y, c["sup"].data[want_idx])
        for r, i in enumerate(want_idx):
            ref = np.array(Image.open(os.path.join(path, "img_align_celeba", c["sup"].index[i])).resize((64, 64)))
            assert np.array_equal(X[r], ref)
    # float32 loaders give exactly uint8 / 255 in fp32 (what the device-side normalisation reproduces bit for bit)
    f32 = G.utils_data.CelebAReader(path, 0.25, 4, split_map=SPLIT).setup_data_loaders(shuffle=False)
    Xf, yf = next(iter(f32["sup"].step()))
    assert Xf.dtype == np.float32 and np.array_equal(Xf, seen[0][0].astype(np.float32) / 255.0) and np.array_equal(yf, seen[0][1])
    # shuffled loaders visit every sample once per pass
    rng = np.random.default_rng(3)
    sh = G.utils_data.CelebAReader(path, 0.25, 3, split_map=SPLIT, dtype="uint8").setup_data_loaders(shuffle=True, rng=rng)
    assert sorted(sh["unsup"].idxs) == list(range(18)) and sh["unsup"].idxs != list(range(18))


def test_prefetching_loader_yields_the_same_batches_in_the_same_order(root):
    path, _ = root
    mk = lambda: G.utils_data.CelebAReader(path, 0.25, 5, split_map=SPLIT, dtype="uint8").setup_data_loaders(
        shuffle=True, rng=np.random.default_rng(7))["unsup"]
    plain, wrapped = mk(), UD.PrefetchLoader(mk(), workers=4, depth=3)
    assert wrapped.n_s == plain.n_s == 18
    a, b = iter(plain.step()), iter(wrapped.step())
    for _ in range(9):                       # 45 samples: wraps around the 18-sample order twice
        (xa, ya), (xb, yb) = next(a), next(b)
        assert xb.dtype == np.uint8 and np.array_equal(xa, xb) and np.array_equal(ya, yb)
    wrapped.close()
    f32 = UD.PrefetchLoader(G.utils_data.CelebAReader(path, 0.25, 5, split_map=SPLIT).setup_data_loaders(
        shuffle=True, rng=np.random.default_rng(7))["unsup"], workers=2, depth=1)
    it = iter(f32.step())
    next(it)
    xf, _ = next(it)
    assert xf.dtype == np.float32 and float(xf.max()) <= 1.0
    f32.close()


def test_wrapped_loaders_resume_like_the_plain_loader_across_step_calls(root, tmp_path):
    """Learner.train() / accuracy() call .step() afresh every epoch: the reference's generator (utils_data.py:77-80) then
    re-yields the batch it read last and reads on.  The wrappers must serve the same sequence over several step() calls
    (and must not drop the batches they decoded ahead)."""
    path, _ = root
    mk = lambda: G.utils_data.CelebAReader(path, 0.25, 4, split_map=SPLIT, dtype="uint8").setup_data_loaders(
        shuffle=True, rng=np.random.default_rng(5))["unsup"]

    def epochs(loader, n_epochs=3, per_epoch=3):
        out = []
        for _ in range(n_epochs):
            it = iter(loader.step())
            out.append([next(it) for _ in range(per_epoch)])
        return out

    want = epochs(mk())
    # the plain loader repeats the last batch of an epoch as the first of the next one
    assert np.array_equal(want[0][-1][0], want[1][0][0]) and not np.array_equal(want[1][0][0], want[1][1][0])
    pre = UD.PrefetchLoader(mk(), workers=3, depth=2)
    cached = UD.CachedLoader(mk(), cache_path=str(tmp_path / "c.npy"), workers=2)
    for got in (epochs(pre), epochs(cached)):
        for ew, eg in zip(want, got):
            for (xa, ya), (xb, yb) in zip(ew, eg):
                assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
    pre.close()

    # reset() rewinds the read position for all three in the same way (the batch read last is still yielded first)
    def with_reset(loader):
        it = iter(loader.step())
        out = [next(it) for _ in range(2)]
        loader.reset()
        it = iter(loader.step())
        return out + [next(it) for _ in range(3)]

    want = with_reset(mk())
    pre = UD.PrefetchLoader(mk(), workers=3, depth=2)
    for got in (with_reset(pre), with_reset(UD.CachedLoader(mk(), cache_path=str(tmp_path / "c.npy")))):
        for (xa, ya), (xb, yb) in zip(want, got):
            assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
    pre.close()


def test_cached_loader_decodes_once_and_serves_identical_batches(root, tmp_path):
    import shutil
    path, _ = root
    work = str(tmp_path / "copy")
    shutil.copytree(path, work)
    mk = lambda: G.utils_data.CelebAReader(work, 0.25, 5, split_map=SPLIT, dtype="uint8").setup_data_loaders(
        shuffle=True, rng=np.random.default_rng(11))["unsup"]
    plain = mk()
    cache = os.path.join(work, "unsup_64x64_u8.npy")
    cached = UD.CachedLoader(mk(), cache_path=cache, workers=3)
    assert os.path.exists(cache) and cached.pixels.shape == (18, 64, 64, 3) and cached.pixels.dtype == np.uint8
    a, b = iter(plain.step()), iter(cached.step())
    seen = []
    for _ in range(9):                       # 45 samples: wraps around the 18-sample order twice
        (xa, ya), (xb, yb) = next(a), next(b)
        assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
        seen.append((xa, ya))
    # a second construction reads the cache: it works with the images gone (only the wrapped loader's own first batch
    # is decoded by its constructor, so those files stay)
    again_src = mk()
    keep = {again_src.cached_data.index[i] for i in again_src.idxs[:5]}
    for f in os.listdir(os.path.join(work, "img_align_celeba")):
        if f not in keep:
            os.remove(os.path.join(work, "img_align_celeba", f))
    again = UD.CachedLoader(again_src, cache_path=cache)
    it = iter(again.step())
    for xa, ya in seen[:6]:
        xd, yd = next(it)
        assert np.array_equal(xa, xd) and np.array_equal(ya, yd)
    np.save(cache, np.zeros((3, 64, 64, 3), np.uint8))
    with pytest.raises(ValueError):
        UD.CachedLoader(again_src, cache_path=cache)
