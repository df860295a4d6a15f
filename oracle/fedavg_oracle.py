"""Oracle: FedAvg of the communicated keys (reference src/train/main.py:50-65).  TEST ONLY."""
from __future__ import annotations

import numpy as np


def get_average_net(client_tensors, nums):
    """main.py:57-64 for one key, in the reference's exact fp32 operation order:
    temp = zeros.float(); for net, num: temp += net[key] * num / total; server.copy_(temp)."""
    total = np.float32(sum(nums))
    temp = np.zeros_like(np.asarray(client_tensors[0], np.float32))
    for x, num in zip(client_tensors, nums):
        temp = temp + (np.asarray(x, np.float32) * np.float32(num)) / total
    return temp
