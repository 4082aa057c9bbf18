"""CPU tests of the MATLAB detection-metric ports (bdpose/detection_metrics.py; reference:
computeAVP.m, computeARP.m, get_R.m, get_v.m, computeGeodesicError.m).  The reference ships no
outputs of these scripts and neither MATLAB nor octave exists in this image ("parity unpinned" for
this row): the checks are hand-worked cases that follow the .m files line by line, identities between
the helpers, and an end-to-end run over annotation / detection files written in the formats the
scripts read."""
import os

import numpy as np
import pytest

import bdpose_oracle as O
from bdpose import detection_metrics as DM


def test_azimuth_bins_follow_the_m_file():
    a = DM.azimuth_intervals(8)
    assert np.allclose(a, [0, 22.5, 67.5, 112.5, 157.5, 202.5, 247.5, 292.5, 337.5])
    f = lambda az: DM.find_interval(az, a)
    assert f(10) == 1 and f(0) == 1          # first half of the wrap-around bin
    assert f(22.5) == 2 and f(67.4) == 2
    assert f(100) == 3 and f(337.4) == 8
    assert f(350) == 1                        # > last edge: wraps to bin 1
    assert f(337.5) == 8                      # == last edge: the loop never breaks -> numel(a) - 1
    assert f(-5) == 0                         # the .m's own quirk for a negative azimuth
    assert np.allclose(DM.azimuth_intervals(4), [0, 45, 135, 225, 315])


def test_rotation_helpers_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(200):
        az, el, ct = rng.uniform(0, 360), rng.uniform(-80, 80), rng.uniform(-60, 60)
        if abs(el) < 1e-3:
            continue
        R = DM.get_R(az, el, ct)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and np.linalg.det(R) > 0
        v = DM.get_v(R)
        assert np.allclose(DM._rodrigues(v), R, atol=1e-9)           # get_v is the log map of Rodrigues
        a2, e2, _ = DM.get_angles(v)
        assert abs(e2 - el) < 1e-6
        assert abs((a2 - az + 180) % 360 - 180) < 1e-6
        assert 0 <= DM.get_azimuth(v) < 360 + 1e-9
    # Euler convention of the python side of the reference: helperFunctions.rotation_matrix == get_R.m
    import helperFunctions as H
    assert np.allclose(H.rotation_matrix(33.0, -12.0, 7.0), DM.get_R(33.0, -12.0, 7.0), atol=1e-12)
    # el == 0 branch of get_angles: azimuth from R(2,1), R(1,1)
    az0, el0, ct0 = DM.get_angles(DM.get_v(DM.get_R(40.0, 0.0, 0.0)))
    assert el0 == 0 and ct0 is None and abs(az0 - 40.0) < 1e-9
    assert np.allclose(DM.get_v(np.eye(3)), 0)


def test_geodesic_error_matches_the_oracle():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((64, 3))
    b = rng.standard_normal((64, 3))
    ref = O.errors_aa(a, b)                                           # axisAngle.get_error's per-pair angle
    got = np.array([DM.geodesic_error(x, y) for x, y in zip(a, b)])
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-7)
    assert DM.geodesic_error(a[0], a[0]) < 1e-3                       # clamp at 1 - 1e-10: acosd -> ~8e-4 deg


def _box(x, y, w=40, h=30):
    return [x, y, x + w, y + h]


def test_score_class_hand_worked_case():
    """One image, two ground-truth objects, three detections: a match with the right view, a duplicate
    of the same object (counts as false), a match with the wrong view."""
    gt_bbox = [np.array([_box(10, 10), _box(200, 50)], dtype=float)]
    v1 = DM.get_v(DM.get_R(30.0, 10.0, 0.0))
    v2 = DM.get_v(DM.get_R(200.0, 20.0, 5.0))
    dets = np.array([_box(12, 11) + [0.9], _box(9, 12) + [0.8], _box(203, 48) + [0.7]], dtype=float)
    far = DM.get_v(DM.get_R(290.0, 20.0, 5.0))
    r = DM.score_class(gt_bbox, [np.stack([v1, v2])], [dets], [np.stack([v1, v1, far])], 'arp')
    assert r['num_total'] == 2 and r['num_correct'] == 2 and r['num_correct_view'] == 1
    assert np.allclose(r['precision'], [1, 0.5, 2 / 3]) and np.allclose(r['recall'], [0.5, 0.5, 1.0])
    assert np.allclose(r['accuracy'], [1, 0.5, 1 / 3])
    assert abs(r['ap'] - (0.5 * 1 + 0.5 * 2 / 3)) < 1e-12
    assert abs(r['aa'] - (0.5 * 1 + 0.5 * 1 / 3)) < 1e-12
    assert r['err'].shape == (2,) and r['err'][0] < 1e-3 and r['err'][1] > 30
    # the same case as azimuth bins: 30 and 200 degrees against predictions at 30, 30 and 290
    q = DM.score_class(gt_bbox, [np.array([30.0, 200.0])], [dets], [np.stack([v1, v1, far])], 'avp', nbins=8)
    assert q['num_correct_view'] == 1 and abs(q['aa'] - r['aa']) < 1e-12
    assert np.allclose(q['err'], [0.0, 90.0], atol=1e-6)
    assert abs(q['med_err'] - 45.0) < 1e-6


def test_score_class_order_ties_and_empty():
    gt = [np.array([_box(0, 0)], dtype=float), np.zeros((0, 4))]
    v = DM.get_v(DM.get_R(10.0, 5.0, 0.0))
    # image 0: a miss listed before the hit with the SAME score (stable sort keeps the miss first);
    # image 1: a detection in an image without objects of the class
    d0 = np.array([_box(300, 300) + [0.5], _box(1, 1) + [0.5]], dtype=float)
    d1 = np.array([_box(5, 5) + [0.9]], dtype=float)
    r = DM.score_class(gt, [np.stack([v]), np.zeros((0, 3))], [d0, d1], [np.stack([v, v]), np.stack([v])], 'arp')
    assert np.allclose(r['precision'], [0, 0, 1 / 3]) and np.allclose(r['recall'], [0, 0, 1])
    assert np.allclose(r['accuracy'], [0, 0, 1 / 3])       # 0 while num_correct == 0 (computeARP.m:128-132)
    assert abs(r['ap'] - 1 / 3) < 1e-12
    # no detections at all: AP 0, median of nothing is NaN
    e = DM.score_class(gt, [np.stack([v]), np.zeros((0, 3))], [np.zeros((0, 5))] * 2, [np.zeros((0, 3))] * 2, 'arp')
    assert e['ap'] == 0 and e['num_total'] == 1 and np.isnan(e['med_err'])
    with pytest.raises(NameError):
        DM.score_class([], [], [], [], 'xyz')


def _write_dataset(root, images):
    """Annotation records, dbinfo.mat and a results file in the formats the .m files read."""
    import scipy.io as spio
    anno = os.path.join(root, 'Annotations')
    dets_path = os.path.join(root, 'dets')
    os.makedirs(dets_path)
    names = np.array([im['name'] for im in images], dtype=object)
    spio.savemat(os.path.join(dets_path, 'dbinfo.mat'), {'image_names': names})
    for im in images:
        objs = np.zeros((len(im['objects']),), dtype=[('class', 'O'), ('difficult', 'O'), ('bbox', 'O'),
                                                      ('viewpoint', 'O')])
        for j, ob in enumerate(im['objects']):
            objs[j]['class'] = ob['cls']
            objs[j]['difficult'] = ob.get('difficult', 0)
            objs[j]['bbox'] = np.asarray(ob['bbox'], dtype=float)
            objs[j]['viewpoint'] = ob['viewpoint']
        for cls in {ob['cls'] for ob in im['objects']}:
            d = os.path.join(anno, '%s_pascal' % cls)
            os.makedirs(d, exist_ok=True)
            spio.savemat(os.path.join(d, im['name'] + '.mat'), {'record': {'objects': objs}})
    n = len(images)
    bbox, ypred, labels = (np.empty((n,), dtype=object) for _ in range(3))
    for i, im in enumerate(images):
        bbox[i] = np.asarray(im['det_boxes'], dtype=float).reshape(-1, 5)
        ypred[i] = np.asarray(im['det_pose'], dtype=float).reshape(-1, 3)
        labels[i] = np.asarray(im['det_labels'], dtype=np.int64).reshape(-1, 1)
    res = os.path.join(root, 'results_dets.mat')
    spio.savemat(res, {'bbox': bbox, 'ypred': ypred, 'labels': labels})
    return anno, dets_path, res


def test_compute_avp_arp_end_to_end(tmp_path, capsys):
    fine = lambda az, el, ct: dict(distance=2.5, azimuth=az, elevation=el, theta=ct, azimuth_coarse=0.0,
                                   elevation_coarse=0.0)
    coarse = lambda az, el: dict(distance=0, azimuth=0.0, elevation=0.0, theta=0.0, azimuth_coarse=az,
                                 elevation_coarse=el)
    car, chair = DM.CLASSES.index('car'), DM.CLASSES.index('chair')
    pose = lambda az, el, ct: DM.get_v(DM.get_R(az, el, ct))
    images = [
        dict(name='img_a',
             objects=[dict(cls='car', bbox=_box(10, 10), viewpoint=fine(40.0, 12.0, 3.0)),
                      dict(cls='car', bbox=_box(150, 20), viewpoint=coarse(180.0, 15.0)),
                      dict(cls='car', bbox=_box(300, 300), viewpoint=fine(10.0, 5.0, 0.0), difficult=1),
                      dict(cls='chair', bbox=_box(60, 200), viewpoint=fine(270.0, 20.0, -4.0))],
             det_boxes=[_box(11, 9) + [0.95], _box(149, 22) + [0.9], _box(61, 199) + [0.8]],
             det_pose=[pose(40.0, 12.0, 3.0), pose(180.0, 15.0, 0.0), pose(270.0, 20.0, -4.0)],
             det_labels=[car, car, chair]),
        dict(name='img_b',
             objects=[dict(cls='chair', bbox=_box(5, 5), viewpoint=fine(100.0, 30.0, 2.0))],
             det_boxes=[_box(6, 6) + [0.7], _box(200, 200) + [0.6]],
             det_pose=[pose(100.0, 30.0, 2.0), pose(0.0, 10.0, 0.0)],
             det_labels=[chair, car]),       # the car detection of img_b: no car annotation file -> skipped
    ]
    anno, dets_path, res = _write_dataset(str(tmp_path), images)
    assert DM.load_image_names(dets_path) == ['img_a', 'img_b']
    ann = DM.load_class_annotation(anno, 'car', 'img_a')
    assert ann['bbox'].shape == (2, 4) and np.allclose(ann['az'], [40.0, 180.0])     # difficult object dropped
    assert DM.load_class_annotation(anno, 'car', 'img_b') is None
    arp = {r['cls']: r for r in DM.computeARP(res, dets_path, anno_path=anno)}
    avp = {r['cls']: r for r in DM.computeAVP(res, 8, dets_path, anno_path=anno)}
    out = capsys.readouterr().out
    assert 'AP = 1.0000' in out and 'car' in out and 'MedErr' in out
    for r in (arp, avp):
        assert r['car']['num_total'] == 2 and r['car']['ap'] == 1.0 and r['car']['aa'] == 1.0
        assert r['chair']['num_total'] == 2 and r['chair']['ap'] == 1.0 and r['chair']['aa'] == 1.0
        assert r['aeroplane']['num_total'] == 0 and np.isnan(r['aeroplane']['med_err'])
    assert arp['car']['med_err'] < 1e-3 and avp['car']['med_err'] < 1e-6
