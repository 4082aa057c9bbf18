"""TEST DOUBLE for the image side of the reference's dataGenerators module (dataGenerators.py:26-77).

The mirror's generator classes subclass whatever `dataGenerators.ImagesAll` is importable (the
reference's own in a deployment: image files, PIL — outside the kernel scope).  The GPU box has no
reference checkout, so the generator tests put this stand-in on sys.path: same constructor, same
attributes (`list_image_names`, `image_names`, `num_images`, `num_classes`, `db_type`, `ydata_type`),
same item keys; images are not read (xdata is zeros), the pose target comes from the oracle's
restatement of helperFunctions.rotation_matrix + axisAngle.get_y / quaternion.get_y."""
import os

import numpy as np
import scipy.io as spio
import torch
from torch.utils.data import Dataset

import bdpose_oracle as O

CLASSES = None       # set by the test (the reference takes helperFunctions.classes)


class ImagesAll(Dataset):
    def __init__(self, db_path, db_type, ydata_type='axis_angle'):
        self.db_path = db_path
        self.classes = list(CLASSES)
        self.num_classes = len(self.classes)
        self.db_type = db_type
        self.ydata_type = ydata_type
        self.list_image_names = []
        for c in self.classes:
            tmp = spio.loadmat(os.path.join(db_path, c + '_info'), squeeze_me=True)
            self.list_image_names.append(np.atleast_1d(tmp['image_names']))
        self.num_images = np.array([len(n) for n in self.list_image_names])
        self.image_names = self.list_image_names

    def __len__(self):
        return np.amax(self.num_images)

    def __getitem__(self, idx):
        xdata, ydata, label = [], [], []
        for i in range(self.num_classes):
            name = str(self.image_names[i][idx % self.num_images[i]])
            parts = name.split('_')
            az, el, ct = float(parts[2][1:]), float(parts[3][1:]), float(parts[4][1:])
            R = O.rotation_matrix(az, el, ct if self.db_type == 'real' else -ct)
            y = O.get_y(R) if self.ydata_type == 'axis_angle' else O.quat_get_y(R)
            ydata.append(torch.from_numpy(np.asarray(y)).float())
            xdata.append(torch.zeros(3, 4, 4))
            label.append(i * torch.ones(1).long())
        return {'xdata': torch.stack(xdata), 'ydata': torch.stack(ydata), 'label': torch.stack(label)}

    def shuffle_images(self):
        self.image_names = [np.random.permutation(self.list_image_names[i]) for i in range(self.num_classes)]
