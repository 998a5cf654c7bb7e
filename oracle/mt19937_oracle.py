"""CPU restatement of the random draws the reference controller makes through torch's CPU generator - TEST
INFRASTRUCTURE ONLY (tests/ and __graft_entry__.smoke() may import it; the product path never does).

The reference draws its workspace samples with ``torch.distributions.Uniform(low, high).sample((N,))``
(franka_test/scripts/control_torch/klerg.py:173,375) = ``low + torch.rand(N, D) * (high - low)`` on the CPU default
generator.  torch is a third-party dependency of the reference (requirements.txt pins torch==2.0.1; the container
has 2.11.0); its CPU generator is the 32-bit Mersenne Twister MT19937 (Matsumoto & Nishimura 1998) and a float32
uniform draw is ``(y & (2**24 - 1)) * 2**-24`` of one tempered output y (ATen/core/TransformationHelper.h
``uniform_real``, ATen/core/MT19937RNGEngine.h).  This file restates that algorithm in numpy on the generator's
serialised state (``torch.get_rng_state()``), and is pinned against ``torch.rand`` itself in
tests/test_rng_oracle.py - the device kernel (csrc/klerg_rng.cu) is then checked against both."""
import struct

import numpy as np

N, M = 624, 397
MATRIX_A, UMASK, LMASK = np.uint32(0x9908B0DF), np.uint32(0x80000000), np.uint32(0x7FFFFFFF)
OFF_LEFT, OFF_NEXT, OFF_STATE = 8, 16, 24  # CPUGeneratorImplStateLegacy: seed u64, left i32, seeded i32, next u64, state u64[624]


def parse_state(blob):
    """bytes of torch.get_rng_state() -> (state uint32[624], left, next)."""
    b = bytes(blob)
    left = struct.unpack_from("<i", b, OFF_LEFT)[0]
    nxt = struct.unpack_from("<Q", b, OFF_NEXT)[0]
    state = np.frombuffer(b, dtype="<u8", count=N, offset=OFF_STATE).astype(np.uint32)
    return state, left, int(nxt)


def pack_state(blob, state, left, nxt):
    """The same serialised generator with its Mersenne-Twister part replaced."""
    b = bytearray(bytes(blob))
    struct.pack_into("<i", b, OFF_LEFT, int(left))
    struct.pack_into("<Q", b, OFF_NEXT, int(nxt))
    b[OFF_STATE:OFF_STATE + 8 * N] = np.asarray(state, dtype=np.uint32).astype("<u8").tobytes()
    return bytes(b)


def _twist(u, v):
    return (((u & UMASK) | (v & LMASK)) >> np.uint32(1)) ^ np.where(v & np.uint32(1), MATRIX_A, np.uint32(0))


def next_block(old):
    """One regeneration of the 624-word state (mt19937_engine::next_state), in the three independent waves the
    recurrence allows."""
    new = np.empty(N, dtype=np.uint32)
    new[:N - M] = old[M:] ^ _twist(old[:N - M], old[1:N - M + 1])
    new[N - M:2 * (N - M)] = new[:N - M] ^ _twist(old[N - M:2 * (N - M)], old[N - M + 1:2 * (N - M) + 1])
    new[2 * (N - M):N - 1] = new[N - M:M - 1] ^ _twist(old[2 * (N - M):N - 1], old[2 * (N - M) + 1:N])
    new[N - 1] = new[M - 1] ^ _twist(old[N - 1:N], new[0:1])[0]
    return new


def temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    return y ^ (y >> np.uint32(18))


def draw_u32(state, left, nxt, count):
    """`count` raw 32-bit outputs -> (values, state, left, next) exactly as `count` calls of the engine."""
    out = np.empty(count, dtype=np.uint32)
    done = 0
    while done < count:
        avail = left - 1  # outputs left in the current block before the engine regenerates
        if avail == 0:
            state, left, nxt = next_block(state), N + 1, 0  # the regenerating call itself reads state[0]
            avail = N
        take = min(avail, count - done)
        out[done:done + take] = temper(state[nxt:nxt + take])
        nxt += take
        left -= take
        done += take
    return out, state, left, nxt


def uniform_samples(blob, n, low, high):
    """(samples [n, D] float32, new generator blob) = Uniform(low, high).sample((n,)) on that generator."""
    low, high = np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32)
    d = low.shape[0]
    state, left, nxt = parse_state(blob)
    raw, state, left, nxt = draw_u32(state, left, nxt, n * d)
    u = (raw & np.uint32((1 << 24) - 1)).astype(np.float32) * np.float32(2.0 ** -24)
    smp = low[None, :] + u.reshape(n, d) * (high - low)[None, :]  # fp32 multiply, then fp32 add (no fused multiply-add)
    return smp.astype(np.float32), pack_state(blob, state, left, nxt)
