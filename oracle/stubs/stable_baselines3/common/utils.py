import random

import numpy as np


def set_random_seed(seed, using_cuda=False):
    random.seed(seed)
    np.random.seed(seed)
