"""The Dataset classes of the label-generation boundary (SURVEY §8b "generator" row) against golden
vectors of the REFERENCE's classes run end to end on a synthetic dataset directory
(tests/golden/make_golden.py generators: dataGenerators.ImagesAll under GBDGenerator / GBDGeneratorQ /
XPBDGeneratorQ / RBDGenerator, objectnetHelperFunctions.TrainImages / TestImages).

The test rebuilds the same directory (names and dictionary come from the fixture), builds the
mirror's classes on it and compares every __getitem__: bins bit-exact, pose targets / residuals /
rotation matrices to 1e-6 absolute (float32 values of fp64 arithmetic whose last-bit rounding of
sin/cos may differ between numpy and the device)."""
import importlib
import os
import pickle
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ATOL = 1e-6


@pytest.fixture()
def dataset(tmp_path, golden, cuda):
    import scipy.io as spio
    from bdpose.kmeans import KMeans
    g = golden("generators")
    classes = [str(c) for c in g["classes"]]
    counts = [int(c) for c in g["counts"]]
    names = [str(n) for n in g["names"]]
    root = str(tmp_path)
    off = 0
    for c, n in zip(classes, counts):
        spio.savemat(os.path.join(root, c + "_info.mat"),
                     {"image_names": np.array(names[off:off + n], dtype=object)})
        off += n
    os.makedirs(os.path.join(root, "data"))
    km = KMeans(n_clusters=16)
    km.cluster_centers_ = np.array(g["centers"])
    dfile = os.path.join(root, "data", "kmeans_dictionary_axis_angle_16.pkl")
    with open(dfile, "wb") as f:
        pickle.dump(km, f)
    return g, root, dfile, classes, counts, names


@pytest.fixture()
def gen_module():
    """binDeltaGenerators of the mirror over the test double of dataGenerators.ImagesAll."""
    fakes = os.path.join(HERE, "fakes")
    sys.path.insert(0, fakes)
    saved = {k: sys.modules.pop(k, None) for k in ("dataGenerators", "binDeltaGenerators")}
    try:
        import dataGenerators
        assert os.path.dirname(os.path.abspath(dataGenerators.__file__)) == fakes
        mod = importlib.import_module("binDeltaGenerators")
        yield mod, dataGenerators
    finally:
        sys.path.remove(fakes)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


@pytest.mark.parametrize("cls,key,kind", [
    ("GBDGenerator", "gbd", "real"), ("GBDGenerator", "gbd", "render"),
    ("GBDGeneratorQ", "gbdq", "real"), ("XPBDGeneratorQ", "xpbdq", "render"),
    ("RBDGenerator", "rbd", "real")])
def test_generator_items_match_reference(dataset, gen_module, cls, key, kind):
    g, root, dfile, classes, counts, names = dataset
    mod, dg = gen_module
    dg.CLASSES = classes
    gen = getattr(mod, cls)(root, kind, dfile)
    assert len(gen) == max(counts)
    assert gen.num_clusters == 16
    k = "%s_%s" % (key, kind)
    for i in range(len(gen)):
        s = gen[i]
        assert set(s) >= {"xdata", "ydata", "label", "ydata_bin", "ydata_res"}
        assert np.array_equal(s["label"].numpy(), g[k + "_label"][i])
        # the test double's pose targets are the reference's (pins the double itself)
        assert np.allclose(s["ydata"].numpy(), g[k + "_ydata"][i], atol=ATOL, rtol=0)
        b = s["ydata_bin"].numpy()
        if key == "xpbdq":
            assert s["ydata_bin"].dtype == torch.float32
            assert np.allclose(b, g[k + "_bin"][i], atol=ATOL, rtol=1e-5)
        else:
            assert s["ydata_bin"].dtype == torch.int64
            assert np.array_equal(b, g[k + "_bin"][i]), "%s item %d: bins differ" % (k, i)
        assert s["ydata_res"].dtype == torch.float32
        assert np.allclose(s["ydata_res"].numpy(), g[k + "_res"][i], atol=ATOL, rtol=0)
        if key == "rbd":
            assert np.allclose(s["ydata_rot"].numpy(), g[k + "_rot"][i], atol=ATOL, rtol=0)
    # shuffle_images(): the labels follow the image NAME (rows are looked up by name), so after a
    # shuffle column c of an item carries the golden labels of wherever that image sat before
    np.random.seed(3)
    gen.shuffle_images()
    s = gen[1]
    off = 0
    for ci, n in enumerate(counts):
        nm = str(gen.image_names[ci][1 % n])
        j = names[off:off + n].index(nm)            # unshuffled item j shows image j of class ci
        off += n
        if key != "xpbdq":
            assert int(s["ydata_bin"][ci]) == int(g[k + "_bin"][j][ci])
        assert np.allclose(s["ydata_res"][ci].numpy(), g[k + "_res"][j][ci], atol=ATOL, rtol=0)


def test_objectnet_datasets_match_reference(dataset, cuda):
    """objectnetHelperFunctions.TrainImages / TestImages (23-107), images included (8x8 PNGs)."""
    PIL = pytest.importorskip("PIL.Image")
    pytest.importorskip("torchvision")
    import objectnetHelperFunctions as OH
    g, root, dfile, classes, counts, names = dataset
    rng = np.random.default_rng(0)
    off = 0
    for c, n in zip(classes[:5], counts[:5]):
        os.makedirs(os.path.join(root, c), exist_ok=True)
        for nm in names[off:off + n]:
            PIL.fromarray(rng.integers(0, 255, (8, 8, 3), dtype=np.uint8)).save(os.path.join(root, c, nm + ".png"))
        off += n
    cwd = os.getcwd()
    os.chdir(root)                      # the reference opens 'data/kmeans_dictionary_...pkl' relatively
    try:
        tr = OH.TrainImages(root, classes[:5], dict_size=16)
        te = OH.TestImages(root, classes[:5], dict_size=16)
    finally:
        os.chdir(cwd)
    assert len(tr) == g["on_train_ydata"].shape[0]
    for i in range(len(tr)):
        s = tr[i]
        assert tuple(s["xdata"].shape) == (5, 3, 224, 224)
        assert np.array_equal(s["label"].numpy(), g["on_train_label"][i])
        assert np.array_equal(s["ydata_bin"].numpy(), g["on_train_ydata_bin"][i])
        assert np.allclose(s["ydata"].numpy(), g["on_train_ydata"][i], atol=ATOL, rtol=0)
        assert np.allclose(s["ydata_res"].numpy(), g["on_train_ydata_res"][i], atol=ATOL, rtol=0)
    assert len(te) == g["on_test_ydata"].shape[0]
    for i in range(len(te)):
        s = te[i]
        assert tuple(s["xdata"].shape) == (3, 224, 224)
        for fld, key in (("label", "on_test_label"), ("ydata_bin", "on_test_ydata_bin")):
            assert s[fld].dtype == torch.int64 and np.array_equal(s[fld].numpy(), g[key][i])
        assert np.allclose(s["ydata"].numpy(), g["on_test_ydata"][i], atol=ATOL, rtol=0)
        assert np.allclose(s["ydata_res"].numpy(), g["on_test_ydata_res"][i], atol=ATOL, rtol=0)
    np.random.seed(0)
    tr.shuffle_images()
    assert tuple(tr[0]["ydata_bin"].shape) == (5,)
