"""Input matrix shared by the oracle tests and the GPU parity tests.

Replicates the shapes the reference's JUnit tests feed the codecs
(JavaFastPFOR/src/test/java/me/lemire/integercompression/*.java),
with seeded numpy data where the reference uses unseeded java.util.Random.
"""
from __future__ import annotations

import numpy as np

ALL_CODECS = list(range(11))
NAMES = ["BP32_VB", "FASTPFOR256_VB", "FASTPFOR128_VB", "IBP32_IVB", "SK_IBP32_IVB", "VB", "SK_BP32_VB", "SK_FASTPFOR256_VB", "SK_FASTPFOR128_VB",
         "XOR128_IVB", "OPTPFD_VB"]
# codecs the batched / thread-per-list kernels take (XorBinaryPacking and OptPFD run in the warp-per-list kernels only)
FAST_CODECS = list(range(9))
FAST_NAMES = NAMES[:9]
DELTA = 0x100


def clustered(rng, n, maxv):
    """ClusteredDataGenerator.java:34-73 in spirit: sorted distinct values with clustered gaps."""
    out = np.empty(n, np.int64)

    def uniform(off, length, lo, hi):
        out[off:off + length] = np.sort(rng.choice(hi - lo, size=length, replace=False) + lo)

    def fill(off, length, lo, hi):
        rng_ = hi - lo
        if rng_ == length or length <= 10:
            uniform(off, length, lo, hi)
            return
        cut = length // 2 + (int(rng.integers(0, rng_ - length - 1)) if rng_ - length - 1 > 0 else 0)
        p = rng.random()
        if p < 0.25:
            uniform(off, length // 2, lo, lo + cut)
            fill(off + length // 2, length - length // 2, lo + cut, hi)
        elif p < 0.5:
            fill(off, length // 2, lo, lo + cut)
            uniform(off + length // 2, length - length // 2, lo + cut, hi)
        else:
            fill(off, length // 2, lo, lo + cut)
            fill(off + length // 2, length - length // 2, lo + cut, hi)

    fill(0, n, 0, maxv)
    return out


def reference_matrix(big: bool = False):
    """-> list of (label, int64 array of Java-int values (may be negative))"""
    rng = np.random.default_rng(0xC0DEC)
    cases = []
    cases.append(("empty", np.zeros(0, np.int64)))
    cases.append(("saul", np.array([2, 3, 4, 5])))                                  # BasicTest.java:47-70
    for n in list(range(1, 130)) + [255, 256, 257, 383, 384, 511, 512, 513, 1023, 1024, 1025]:
        cases.append((f"iota{n}", np.arange(n)))                                        # varyingLengthTest :74-103, BoundaryTest
    for n in range(128, 4097, 128 * 3):
        cases.append((f"iota{n}", np.arange(n)))
    for n in list(range(128, 140)) + [256, 512, 1024, 4096]:
        d = np.arange(n)
        d[127] = -1
        cases.append((f"minus1at127_{n}", d))                                           # varyingLengthTest2 :108-148
    for n in (133, 1026):                                                               # testUnsortedExample :502-622
        d = np.full(n, 3)
        d[::5] = 100
        d[::533] = 10000
        d[5] = -311
        cases.append((f"unsorted{n}", d))
    for pos in (5, 127):
        d = np.zeros(128, np.int64)
        d[pos] = -1
        cases.append((f"lone_minus1_{pos}", d))
    d = np.zeros(256, np.int64)
    d[126] = -1
    cases.append(("fastpfor_b0_one_exception", d))                                      # fastPforTest :627-664
    cases.append(("adhoc_0_16384", np.array([0, 16384])))                               # AdhocTest.java:17-46
    cases.append(("adhoc_65535", np.array([65535, 65535])))
    cases.append(("adhoc_minus1", np.array([-1])))
    for n in (0, 1, 31, 32, 33, 100, 127, 128, 129, 1000, 4095, 4096):                   # SkippableBasicTest consistentTest
        cases.append((f"kmod128_{n}", np.arange(n) % 128))
    for n in (1, 10, 100, 1000, 10000):                                                 # IntCompressorTest :32-76
        cases.append((f"3k+5_{n}", 3 * np.arange(n) + 5))
    for b in range(0, 33):                                                              # all widths, random payloads
        hi = (1 << b)
        v = rng.integers(0, hi, size=256 + 32 + 7, dtype=np.uint64).astype(np.int64)
        v = np.where(v >= 2**31, v - 2**32, v)
        cases.append((f"width{b}", v))
    for n, sp in ((1 << 10, 1), (1 << 10, 9), (1 << 14, 5), (1 << 14, 13)):              # basictest clustered data
        cases.append((f"clustered_{n}_{sp}", clustered(rng, n, n << sp)))
    for n in (300, 2000, 70000):                                                        # exception-heavy blocks
        v = rng.integers(0, 16, size=n)
        mask = rng.random(n) < 0.07
        v = np.where(mask, rng.integers(0, 1 << 30, size=n), v)
        cases.append((f"exceptions_{n}", v))
    v = rng.integers(0, 4, size=2048)
    v[rng.random(2048) < 0.1] += 4  # maxbits - b == 1 streams (implicit exception values)
    cases.append(("index1_exceptions", v))
    cases.append(("gaps100", np.cumsum(rng.geometric(0.01, size=1000))))                # breach-list-like (1% rate)
    cases.append(("gaps100_short", np.cumsum(rng.geometric(0.01, size=101))))
    cases.append(("full_range", rng.integers(-2**31, 2**31, size=777)))
    cases.append(("vb_5bytes", np.array([2**28, 2**28 - 1, 2**31 - 1, -1, -2**31, 127, 128, 16383, 16384, 2**21 - 1, 2**21])))
    if big:
        d = np.full(1333333, 3)                                                         # multi-page FastPFOR (decode only on GPU)
        d[::5] = 100
        d[::533] = 10000
        d[5] = -311
        cases.append(("unsorted1333333", d))
        cases.append(("sorted2342351", 3 * np.arange(2342351, dtype=np.int64) + 5))
    return cases


def to_i32(a) -> np.ndarray:
    return (np.asarray(a, np.int64) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


def breach_like_lists(rng, nlists: int, n_trials: int, rate: float):
    """CSR batch of ascending trial-index lists (what stage 3 hands the codec)."""
    lists = [np.flatnonzero(rng.random(n_trials) < rate).astype(np.int64) for _ in range(nlists)]
    off = np.zeros(nlists + 1, np.int64)
    np.cumsum([len(x) for x in lists], out=off[1:])
    vals = np.concatenate(lists) if off[-1] else np.zeros(0, np.int64)
    return to_i32(vals), off
