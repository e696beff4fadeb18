"""The in-place scratch of the one-CTA-per-spectrum kernel (csrc/solo65536.cu, Geo<M>), restated in Python and
checked exhaustively: the two block layouts are bijections, and in either direction the blocks the outer stages of
iteration i write are exactly the blocks of the rows the inner transforms of iteration i have consumed -- for every
outer radix, for the plain transpose used at M = 16 and for the tensor-memory cells used at M <= 4.  (CPU test of
the index arithmetic only; the kernel itself is tested on the GPU against the two-kernel path and the oracle.)"""
import pytest


def geo(M, transpose16=True):
    AL, C, IPS = 32 // M, 16 // M, M // 2
    t16 = transpose16 and M == 16
    blk0 = lambda ko, a: ko * 16 + a
    g1 = (lambda ko: ko) if t16 else (lambda ko: (ko & 1) * 16 + (ko >> 1) * AL)
    f1 = (lambda a: a * 16) if t16 else (lambda a: (a // AL) * 32 + a % AL)
    blk1 = lambda ko, a: g1(ko) + f1(a)
    return AL, C, IPS, blk0, blk1, g1, f1


@pytest.mark.parametrize("M", [2, 4, 8, 16])
@pytest.mark.parametrize("transpose16", [True, False])
def test_layouts_are_bijections_and_reuse_in_place(M, transpose16):
    AL, C, IPS, blk0, blk1, g1, f1 = geo(M, transpose16)
    cells = M * 16
    for blk in (blk0, blk1):
        assert sorted(blk(ko, a) for ko in range(M) for a in range(16)) == list(range(cells))
    for i in range(IPS):
        produced_a = [i * AL + half * C + c for half in (0, 1) for c in range(C)]      # a values of the two outer stages
        assert sorted(produced_a) == list(range(i * AL, (i + 1) * AL))
        for cur, nxt in ((blk0, blk1), (blk1, blk0)):
            consumed = {cur(ko, a) for ko in (2 * i, 2 * i + 1) for a in range(16)}      # rows read for iteration i
            produced = {nxt(ko, a) for ko in range(M) for a in produced_a}               # written during iteration i
            assert consumed == produced
    # what the kernel relies on when it splits an address into a run-time base and a compile-time constant
    for i in range(IPS):
        for half in (0, 1):
            for c in range(C):
                a = i * AL + half * C + c
                assert f1(a) == f1(i * AL + half * C) + f1(c)
    # the helper warps' walk over a row in layout 1: runs of kRun blocks, kRunStride1 blocks apart
    run, stride = (1, 16) if (transpose16 and M == 16) else (AL, 32)
    for ko in range(M):
        assert [blk1(ko, a) - g1(ko) for a in range(16)] == [(a // run) * stride + a % run for a in range(16)]


def test_tensor_memory_budget():
    """M <= 4: accumulators (16 M columns from 0) and hand-off cells (2 columns each from column 64) fit a thread's 256
    columns without overlap; M = 8 would not."""
    for M in (2, 4):
        assert 16 * M <= 64 and 64 + 2 * 16 * M <= 256
    assert 16 * 8 + 2 * 16 * 8 > 256
