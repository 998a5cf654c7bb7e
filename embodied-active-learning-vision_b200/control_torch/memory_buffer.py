"""Trajectory memory buffer (reference control_torch/memory_buffer.py:38-92) with a device mirror.

Indices are drawn on the host with ``torch.randperm`` in the reference's RNG
order (bit-exact selection); rows are gathered on the GPU by klerg_gather_rows.
"""
import random

import torch

from . import _cabi as cabi
from . import engine


class MemoryBuffer(object):
    """List-based buffer kept for API parity (reference :6-36); not used by Robot."""

    def __init__(self, capacity):
        self.capacity = capacity
        self.buffer = []
        self.position = 0

    def push(self, state):
        if len(self.buffer) < self.capacity:
            self.buffer.append(None)
        self.buffer[self.position] = state
        self.position = (self.position + 1) % self.capacity

    def sample(self, batch_size):
        return random.sample(self.buffer, batch_size)

    def get_recent(self, batch_size):
        if len(self.buffer) <= batch_size:
            return self.buffer[-batch_size:]
        if self.position - batch_size < 0:
            return self.buffer[: self.position] + self.buffer[(self.position - batch_size):]
        return self.buffer[(self.position - batch_size): self.position]

    def __len__(self):
        return len(self.buffer)

    def reset(self):
        self.position = 0
        self.buffer = []


class MemoryBuffer_torch(torch.nn.Module):
    """Ring buffer [capacity, state_dim]; host copy for the API, fp32 device copy for the kernels."""

    def __init__(self, capacity, state_dim, dtype=float):
        super().__init__()
        cabi.require_cuda()
        self.capacity = capacity
        self.position = 0
        self.full_buffer = False
        self.buffer = torch.empty((capacity, state_dim), dtype=dtype)
        self.device_buffer = torch.zeros((capacity, state_dim), dtype=torch.float32, device="cuda")

    def push(self, state):
        if (self.position + 1) == self.capacity:
            self.full_buffer = True
        row = torch.as_tensor(state).detach()
        self.buffer[self.position] = row.to(self.buffer.dtype).cpu()
        self.device_buffer[self.position].copy_(row.to(torch.float32), non_blocking=True)
        self.position = (self.position + 1) % self.capacity

    def draw_indices(self, batch_size):
        """Host-side index draw, same RNG consumption as the reference's sample() (:52-63)."""
        n = len(self)
        if n == 0:
            return torch.empty(0, dtype=torch.int64)
        return torch.randperm(n)[: min(batch_size, n)]

    def sample_device(self, batch_size):
        """(rows [M, S] on the GPU, indices [M] on the host)."""
        idx = self.draw_indices(batch_size)
        rows = engine.gather_rows(self.device_buffer, idx.to("cuda", non_blocking=True))
        return rows, idx

    def partition_device(self, idx):
        """All buffer rows on the GPU with the drawn ones first: (rows [len, S], number of drawn rows).

        Both groups keep the buffer's time order (consecutive states of the robot's path are close together, which
        keeps the staged chunks of the pair pass compact); the order inside a group does not matter to the sums."""
        n = len(self)
        m = int(idx.numel())
        if m >= n:
            return self.get_all_device(), n
        drawn = torch.zeros(n, dtype=torch.bool)
        drawn[idx] = True
        order = torch.cat([drawn.nonzero().flatten(), (~drawn).nonzero().flatten()])
        return engine.gather_rows(self.device_buffer, order.to("cuda", non_blocking=True)), m

    def sample(self, batch_size):
        rows, _ = self.sample_device(batch_size)
        return rows.to(device="cpu", dtype=self.buffer.dtype)

    def get_recent(self, batch_size):
        if self.position > batch_size:
            return self.buffer[self.position - batch_size: self.position].clone()
        if self.full_buffer:
            return torch.vstack([self.buffer[: self.position], self.buffer[self.position - batch_size:]])
        return self.buffer[: self.position].clone()

    def get_all(self):
        return self.buffer.clone() if self.full_buffer else self.buffer[: self.position].clone()

    def get_all_device(self):
        return self.device_buffer if self.full_buffer else self.device_buffer[: self.position]

    def __len__(self):
        return self.capacity if self.full_buffer else self.position

    def seed(self, seed):
        torch.manual_seed(seed)

    def reset(self):
        self.position = 0
        self.full_buffer = False


class AvoidDist(torch.nn.Module):
    """Imported by the reference's klerg.py but never instantiated (memory_buffer.py:95-147)."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("AvoidDist is dead code in the reference and is not ported")
