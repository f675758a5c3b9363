"""BASELINE config 2 at FULL size (|R| = 2^24 build x |S| = 2^28 probe, unique
permutation keys) on device-resident synthetic columns, checked through
size-independent properties that need no oracle run:
  * every R key has exactly one partner        => matches == 2^24
  * SUM(R.c1) over the result                  == sum of the whole R payload column
  * SUM(S.c1) over the result                  == sum of S payloads of rows whose key < 2^24
  * swapping the operands changes nothing
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config2_full_size_properties(gpu, orc):
    kr_bits, ks_bits = 24, 28
    nr, ns = 1 << kr_bits, 1 << ks_bits
    r0, r1, s0, s1 = (gpu.DeviceColumn(n) for n in (nr, nr, ns, ns))
    gpu.synth_column_device(r0.ptr, 0, nr, gpu.SYNTH_PERM, kr_bits, gpu.SEED_R)
    gpu.synth_column_device(r1.ptr, 0, nr, gpu.SYNTH_PAYLOAD, 0, gpu.SEED_R + 1)
    gpu.synth_column_device(s0.ptr, 0, ns, gpu.SYNTH_PERM, ks_bits, gpu.SEED_S)
    gpu.synth_column_device(s1.ptr, 0, ns, gpu.SYNTH_PAYLOAD, 0, gpu.SEED_S + 1)
    # the device generator equals the CPU generator (spot checks)
    assert np.array_equal(s0.to_host(0, 1 << 16), orc.synth_column(1 << 16, 0, ks_bits, gpu.SEED_S))
    assert np.array_equal(s1.to_host(ns - 1000, 1000), orc.synth_column(1000, 1, 0, gpu.SEED_S + 1, first=ns - 1000))
    assert np.array_equal(r0.to_host(12345, 4096), orc.synth_column(4096, 0, kr_bits, gpu.SEED_R, first=12345))
    # expected sums on the CPU, chunk by chunk, from the generator alone
    want_r = int(orc.synth_column(nr, 1, 0, gpu.SEED_R + 1).sum(dtype=np.uint64))
    want_s, chunk = 0, 1 << 24
    for first in range(0, ns, chunk):
        keys = orc.synth_column(chunk, 0, ks_bits, gpu.SEED_S, first=first)
        pay = orc.synth_column(chunk, 1, 0, gpu.SEED_S + 1, first=first)
        want_s += int(pay[keys < nr].sum(dtype=np.uint64))
    sums, m = gpu.join_sum_device(r0.ptr, nr, s0.ptr, ns, [r1.ptr, s1.ptr], [0, 1], (1 << ks_bits) - 1)
    assert m == nr
    assert sums == [want_r, want_s % (1 << 64)]
    sums2, m2 = gpu.join_sum_device(s0.ptr, ns, r0.ptr, nr, [r1.ptr, s1.ptr], [1, 0], (1 << ks_bits) - 1)
    assert m2 == nr and sums2 == sums
    for c in (r0, r1, s0, s1):
        c.free()
