"""GPU version of the reference's learnKmeansDictionary.py (learnKmeansDictionary.py:1-53):

    python -m bdpose.learn_dictionary K [--image_path data/renderforcnn] [--out FILE] [--synthetic N]

Same steps — collect the rendered images' names, turn every name into its axis-angle pose target
(parse_name -> rotation_matrix(az, el, -ct) -> get_y), fit a K-key dictionary, pickle an estimator with
`.n_clusters`, `.cluster_centers_`, `.predict` to data/kmeans_dictionary_axis_angle_K.pkl — with the
targets computed in one bdp_euler_to_pose launch and the fit on bdpose.kmeans.KMeans (key-grid Lloyd).
`--synthetic N` fits on N uniformly random rotations instead of a dataset (no images needed).
"""
import argparse
import pickle

import numpy as np
import torch

from . import ops
from .kmeans import KMeans


def pose_targets_from_names(image_names, sign_ct=-1.0):
    """[N] rendered-image names -> [N,3] float64 axis-angle targets (learnKmeansDictionary.py:31-37)."""
    from helperFunctions import parse_name      # the reference's own parser (string handling)
    eul = np.zeros((len(image_names), 3))
    for i, name in enumerate(image_names):
        _, _, az, el, ct, _ = parse_name(name)
        eul[i] = (az, el, sign_ct * ct)
    aa, _ = ops.euler_to_pose(torch.from_numpy(eul).cuda(), want_aa=True, want_quat=False)
    return aa.cpu().numpy()


def synthetic_targets(n, seed=0):
    rng = np.random.default_rng(seed)
    eul = np.stack([rng.uniform(0, 360, n), rng.uniform(-90, 90, n), rng.uniform(-180, 180, n)], 1)
    aa, _ = ops.euler_to_pose(torch.from_numpy(eul).cuda(), want_aa=True, want_quat=False)
    return aa.cpu().numpy()


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("num_clusters", type=int)
    ap.add_argument("--image_path", default="data/renderforcnn")
    ap.add_argument("--out", default=None)
    ap.add_argument("--synthetic", type=int, default=0)
    ap.add_argument("--n_init", type=int, default=10)      # the sklearn default the reference ran with
    args = ap.parse_args(argv)
    K = args.num_clusters
    print('num_clusters: ', K)
    out = args.out or 'data/kmeans_dictionary_axis_angle_' + str(K) + '.pkl'
    if args.synthetic:
        ydata = synthetic_targets(args.synthetic)
    else:
        from dataGenerators import ImagesAll
        train_data = ImagesAll(args.image_path, 'render')
        ydata = pose_targets_from_names(np.concatenate(train_data.list_image_names))
    print('\nData size: ', ydata.shape)
    kmeans = KMeans(K, verbose=1, n_init=args.n_init, n_jobs=10)
    kmeans.fit(ydata)
    print(kmeans.cluster_centers_)
    with open(out, 'wb') as fid:
        pickle.dump(kmeans, fid)
    with open(out, 'rb') as fid:                              # load and check, as the reference does
        print(pickle.load(fid).cluster_centers_)
    return out


if __name__ == "__main__":
    main()
