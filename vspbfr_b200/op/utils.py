"""The few filesystem / seeding helpers the reference scripts import from ``op.utils``
(/root/reference/op/utils.py; used by restoration_test.py:9,17).  Image-mask helpers of
that file are data-pipeline code and out of scope."""
import os
import random
import shutil

import numpy as np
import torch


def mkdirs(path):
    os.makedirs(path, exist_ok=True)


def delete_dirs(path):
    if os.path.isdir(path):
        shutil.rmtree(path)


def set_random_seed(seed, deterministic=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
