import math
import numpy as np


def normalized_vector(vec):
    vec = np.asarray(vec, dtype=float)
    return vec / math.sqrt((vec ** 2).sum())
