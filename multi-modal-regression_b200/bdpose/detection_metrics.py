"""Detection metrics of the reference's MATLAB evaluation: numpy ports of
`computeAVP.m` (azimuth-bin viewpoint precision) and `computeARP.m` (rotation precision at 30 degrees)
with their helpers `find_interval`, `get_angles` (computeAVP.m:152-178), `get_R.m`, `get_v.m` and
`computeGeodesicError.m`; `VOCap` / `box_overlap` live in bdpose.metrics.  They consume what
evaluateModelDetectedBBoxes.py:175-189 writes (`results/*_dets.mat` with `bbox`, `ypred`, `labels`
per image) and the Pascal3D+ annotation records.

Host code (numpy / scipy.io): the work is a walk over a few thousand annotation files and a sort per
class — there is no device path and none is claimed.  The scoring core takes plain arrays
(`score_class`), so it can be used and tested without the dataset on disk.

MATLAB semantics kept on purpose: 1-based bin indices, `max` returns the FIRST maximum, `sort(...,
'descend')` is stable, the quirks of `find_interval` at the interval ends, `median([])` = NaN.
"""
import os

import numpy as np

from .metrics import VOCap, box_overlap

CLASSES = ('aeroplane', 'bicycle', 'boat', 'bottle', 'bus', 'car', 'chair', 'diningtable',
           'motorbike', 'sofa', 'train', 'tvmonitor')          # computeAVP.m:13-14


# ---- helpers ------------------------------------------------------------------------------------------
def azimuth_intervals(nbins):
    """computeAVP.m:5 — [0, w/2 : w : 360 - w/2] with w = 360 / nbins."""
    w = 360.0 / nbins
    return np.concatenate([[0.0], w / 2 + w * np.arange(nbins)])


def find_interval(azimuth, a):
    """computeAVP.m:167-178 (1-based bin; bin 1 wraps around 0/360).  As in the .m: an azimuth equal
    to the last edge lands in the last bin, a negative one in bin 0."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    i = a.size                                   # value of the loop variable if the loop never breaks
    for k in range(a.size):
        if azimuth < a[k]:
            i = k + 1
            break
    ind = i - 1
    if azimuth > a[-1]:
        ind = 1
    return ind


def _skew(v):
    # reshape(v * proj, [3, 3]) with MATLAB's column-major reshape (computeAVP.m:155-158)
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def _rodrigues(y, eps=1e-10):
    y = np.asarray(y, dtype=np.float64).reshape(3)
    t = np.linalg.norm(y)
    sv = _skew(y / max(t, eps))
    return np.eye(3) + np.sin(t) * sv + (1 - np.cos(t)) * (sv @ sv)


def get_R(az, el, ct):
    """get_R.m: Rz(ct) * Rx(el) * Rz(az), angles in degrees."""
    a, b, c = np.deg2rad([az, el, ct])
    Ra = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    Rb = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    Rc = np.array([[np.cos(c), -np.sin(c), 0], [np.sin(c), np.cos(c), 0], [0, 0, 1]])
    return Rc @ Rb @ Ra


def get_v(R):
    """get_v.m: axis-angle vector of a rotation matrix (zero axis when R is symmetric)."""
    R = np.asarray(R, dtype=np.float64)
    theta = np.arccos(min(1.0, max(0.5 * (np.trace(R) - 1), -1.0)))
    tmp = 0.5 * (R - R.T)
    y = np.array([tmp[2, 1], tmp[0, 2], tmp[1, 0]])
    n = np.linalg.norm(y)
    u = y / n if n else np.zeros(3)
    return theta * u


def get_angles(y):
    """computeAVP.m:152-165 — (az, el, ct) in degrees of the rotation exp([y]x).  `ct` is None on the
    el == 0 branch (the .m leaves it unset there).  acosd's argument is clipped to [-1, 1] (MATLAB
    would return a complex number for a trace rounded past 1)."""
    R = _rodrigues(y)
    el = np.sign(-R[1, 2]) * np.degrees(np.arccos(np.clip(R[2, 2], -1.0, 1.0)))
    ct = None
    if el != 0:
        s = np.sin(np.radians(el))
        az = np.degrees(np.arctan2(R[2, 0] / s, R[2, 1] / s))
        with np.errstate(divide='ignore', invalid='ignore'):
            ct = np.degrees(np.arctan(-R[0, 2] / R[1, 2]))
    else:
        az = np.degrees(np.arctan2(R[1, 0], R[0, 0]))
    return az, el, ct


def get_azimuth(y):
    """computeAVP.m:147-150: azimuth in [0, 360)."""
    az = get_angles(y)[0]
    return az + 360 if az < 0 else az


def geodesic_error(v1, v2, eps=1e-10):
    """computeGeodesicError.m for one pair of axis-angle vectors, degrees."""
    R = _rodrigues(v1, eps).T @ _rodrigues(v2, eps)
    tmp = min(1 - eps, max(-1 + eps, 0.5 * (np.trace(R) - 1)))
    return abs(np.degrees(np.arccos(tmp)))


# ---- scoring core -------------------------------------------------------------------------------------
def score_class(gt_bbox, gt_view, det_boxes, det_view, mode, nbins=None):
    """The per-class body of computeAVP.m:27-146 / computeARP.m:28-155 on arrays.

    gt_bbox[i]  [n_i, 4]  ground-truth boxes of the class in image i (non-difficult objects)
    gt_view[i]  'avp': [n_i] azimuths in degrees; 'arp': [n_i, 3] axis-angle vectors
    det_boxes[i] [m_i, 5] detections of the class in image i (x1 y1 x2 y2 score), in file order
    det_view[i]  [m_i, 3] predicted axis-angle pose of every detection
    Images without an annotation file for the class are simply left out of the lists (the .m skips
    them, detections included).

    Returns dict(ap, aa, med_err, num_total, num_correct, num_correct_view, recall, precision,
    accuracy, err)."""
    if mode not in ('avp', 'arp'):
        raise NameError('Unknown mode passed')
    if mode == 'avp':
        edges = azimuth_intervals(nbins)
    energy, correct, correct_view, err = [], [], [], []
    total = 0
    for bbox, view, dets, ypred in zip(gt_bbox, gt_view, det_boxes, det_view):
        bbox = np.asarray(bbox, dtype=np.float64).reshape(-1, 4)
        n = bbox.shape[0]
        total += n
        taken = np.zeros(n, dtype=bool)
        dets = np.asarray(dets, dtype=np.float64).reshape(-1, 5)
        ypred = np.asarray(ypred, dtype=np.float64).reshape(-1, 3)
        if mode == 'avp':
            view = np.asarray(view, dtype=np.float64).reshape(-1)
            az_gt = [find_interval(a, edges) for a in view]
        else:
            view = np.asarray(view, dtype=np.float64).reshape(-1, 3)
        for j in range(dets.shape[0]):
            energy.append(dets[j, 4])
            ok = ok_view = 0
            if n > 0:
                o = box_overlap(bbox, dets[j, :4])
                index = int(np.argmax(o))                 # first maximum, like MATLAB's max
                if o[index] >= 0.5 and not taken[index]:
                    ok = 1
                    taken[index] = True
                    if mode == 'avp':
                        az_pred = get_azimuth(ypred[j])
                        err.append(abs(az_pred - view[index]))
                        ok_view = 1 if find_interval(az_pred, edges) == az_gt[index] else 0
                    else:
                        theta = geodesic_error(view[index], ypred[j])
                        err.append(theta)
                        ok_view = 1 if theta < 30 else 0
            correct.append(ok)
            correct_view.append(ok_view)
    energy = np.asarray(energy, dtype=np.float64)
    order = np.argsort(-energy, kind='stable')            # sort(energy, 'descend') keeps ties in order
    correct = np.asarray(correct, dtype=np.float64)[order]
    correct_view = np.asarray(correct_view, dtype=np.float64)[order]
    n = energy.size
    num_positive = np.arange(1, n + 1, dtype=np.float64)
    cum_correct = np.cumsum(correct)
    cum_view = np.cumsum(correct_view)
    precision = cum_correct / num_positive if n else np.zeros(0)
    accuracy = np.where(cum_correct != 0, cum_view / num_positive, 0.0) if n else np.zeros(0)
    with np.errstate(divide='ignore', invalid='ignore'):
        recall = cum_correct / float(total) if n else np.zeros(0)
    return dict(ap=VOCap(recall, precision), aa=VOCap(recall, accuracy),
                med_err=float(np.median(err)) if err else float('nan'),
                num_total=int(total), num_correct=int(cum_correct[-1]) if n else 0,
                num_correct_view=int(cum_view[-1]) if n else 0,
                recall=recall, precision=precision, accuracy=accuracy, err=np.asarray(err))


# ---- files --------------------------------------------------------------------------------------------
def _cell(x):
    """A MATLAB cell / savemat'ed python list -> list of arrays."""
    x = np.asarray(x)
    if x.dtype == object:
        return [np.asarray(v) for v in x.reshape(-1)]
    return [x[i] for i in range(x.shape[0])]


def load_detections(filename):
    """`results/*_dets.mat` as written by evaluateModelDetectedBBoxes.py:177: per-image `bbox` [m,5],
    `ypred` [m,3], `labels` [m] (0-based class ids)."""
    import scipy.io as spio
    tmp = spio.loadmat(filename)
    boxes = [np.asarray(b, dtype=np.float64).reshape(-1, 5) for b in _cell(tmp['bbox'])]
    ypred = [np.asarray(y, dtype=np.float64).reshape(-1, 3) for y in _cell(tmp['ypred'])]
    labels = [np.asarray(l).reshape(-1).astype(np.int64) for l in _cell(tmp['labels'])]
    return boxes, ypred, labels


def load_image_names(dets_path):
    """`<dets_path>/dbinfo.mat` -> list of image names (computeAVP.m:17-19)."""
    import scipy.io as spio
    tmp = spio.loadmat(os.path.join(dets_path, 'dbinfo'), squeeze_me=True)
    return [str(s).strip() for s in np.atleast_1d(tmp['image_names'])]


def load_class_annotation(anno_path, cls, image_name):
    """The non-difficult objects of class `cls` in one Pascal3D+ annotation record
    (computeAVP.m:38-62, computeARP.m:41-70): None when the file does not exist, else
    dict(bbox [n,4], az [n], view [n,3])."""
    import scipy.io as spio
    fn = os.path.join(anno_path, '%s_pascal' % cls, image_name + '.mat')
    if not os.path.exists(fn):
        return None
    rec = spio.loadmat(fn, squeeze_me=True, struct_as_record=False)['record']
    bbox, az_l, view = [], [], []
    for ob in np.atleast_1d(rec.objects):
        if str(getattr(ob, 'class')) != cls or int(ob.difficult):
            continue
        vp = ob.viewpoint
        if float(vp.distance) == 0:
            az, el = float(vp.azimuth_coarse), float(vp.elevation_coarse)
        else:
            az, el = float(vp.azimuth), float(vp.elevation)
        ct = float(vp.theta)
        bbox.append(np.asarray(ob.bbox, dtype=np.float64).reshape(4))
        az_l.append(az)
        view.append(get_v(get_R(az, el, ct)))
    return dict(bbox=np.asarray(bbox, dtype=np.float64).reshape(-1, 4), az=np.asarray(az_l, dtype=np.float64),
                view=np.asarray(view, dtype=np.float64).reshape(-1, 3))


def _compute(filename, dets_path, anno_path, mode, nbins, classes, verbose):
    image_names = load_image_names(dets_path)
    boxes_all, ypred_all, labels_all = load_detections(filename)
    out = []
    for cls_id, cls in enumerate(classes):
        gt_bbox, gt_view, det_boxes, det_view = [], [], [], []
        for i, name in enumerate(image_names):
            ann = load_class_annotation(anno_path, cls, name)
            if ann is None:
                continue
            ind = np.nonzero(labels_all[i] == cls_id)[0]
            gt_bbox.append(ann['bbox'])
            gt_view.append(ann['az'] if mode == 'avp' else ann['view'])
            det_boxes.append(boxes_all[i][ind])
            det_view.append(ypred_all[i][ind])
        r = score_class(gt_bbox, gt_view, det_boxes, det_view, mode, nbins)
        r['cls'] = cls
        out.append(r)
        if verbose:
            print(cls)
            print('AP = %.4f' % r['ap'])
            print('AA = %.4f' % r['aa'])
            if mode == 'avp':
                print('MedErr = %.4f' % r['med_err'])
            else:
                tot = max(r['num_total'], 1)
                print('Stats: \t num_total=%d \t percent_correct=%0.2f \t percent_correct_view:%0.2f \t '
                      'MedErr = %2.1f ' % (r['num_total'], r['num_correct'] / tot,
                                           r['num_correct_view'] / tot, r['med_err']))
    return out


def computeAVP(filename, nbins, dets_path, anno_path=os.path.join('data', 'pascal3d', 'Annotations'),
               classes=CLASSES, verbose=True):
    """computeAVP.m: per class AP, AVP ("AA") over `nbins` azimuth bins and the median azimuth error
    of the matched detections.  Returns a list of per-class dicts (see score_class)."""
    return _compute(filename, dets_path, anno_path, 'avp', nbins, classes, verbose)


def computeARP(filename, dets_path, anno_path=os.path.join('data', 'pascal3d', 'Annotations'),
               classes=CLASSES, verbose=True):
    """computeARP.m: per class AP, ARP ("AA": geodesic error below 30 degrees) and the matching
    statistics.  `filename` is the full path of the results file (the .m prefixes 'results/')."""
    return _compute(filename, dets_path, anno_path, 'arp', None, classes, verbose)
