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
    torch = pytest.importorskip("torch")
    kr_bits, ks_bits = 24, 28
    nr, ns = 1 << kr_bits, 1 << ks_bits
    dev = torch.device("cuda:0")
    cols = {name: torch.empty(n, dtype=torch.int64, device=dev) for name, n in
            [("r0", nr), ("r1", nr), ("s0", ns), ("s1", ns)]}
    gpu.synth_column_device(cols["r0"].data_ptr(), 0, nr, gpu.SYNTH_PERM, kr_bits, gpu.SEED_R)
    gpu.synth_column_device(cols["r1"].data_ptr(), 0, nr, gpu.SYNTH_PAYLOAD, 0, gpu.SEED_R + 1)
    gpu.synth_column_device(cols["s0"].data_ptr(), 0, ns, gpu.SYNTH_PERM, ks_bits, gpu.SEED_S)
    gpu.synth_column_device(cols["s1"].data_ptr(), 0, ns, gpu.SYNTH_PAYLOAD, 0, gpu.SEED_S + 1)
    torch.cuda.synchronize()
    # the device generator equals the CPU generator (spot check, first 2^16 rows and a tail slice)
    assert np.array_equal(cols["s0"][: 1 << 16].cpu().numpy().view(np.uint64),
                          orc.synth_column(1 << 16, 0, ks_bits, gpu.SEED_S))
    assert np.array_equal(cols["s1"][ns - 1000:].cpu().numpy().view(np.uint64),
                          orc.synth_column(1000, 1, 0, gpu.SEED_S + 1, first=ns - 1000))
    want_r = int(cols["r1"].sum().item())                       # < 2^48, no wrap
    want_s = int(cols["s1"][cols["s0"] < nr].sum().item())
    sums, m = gpu.join_sum_device(cols["r0"].data_ptr(), nr, cols["s0"].data_ptr(), ns,
                                  [cols["r1"].data_ptr(), cols["s1"].data_ptr()], [0, 1], (1 << ks_bits) - 1)
    assert m == nr
    assert sums == [want_r, want_s]
    sums2, m2 = gpu.join_sum_device(cols["s0"].data_ptr(), ns, cols["r0"].data_ptr(), nr,
                                    [cols["r1"].data_ptr(), cols["s1"].data_ptr()], [1, 0], (1 << ks_bits) - 1)
    assert m2 == nr and sums2 == sums
