"""Multi-GPU path on real GPUs (needs >= 2 devices; skipped otherwise): one process per GPU, spectra
sharded by index, one NCCL allreduce of pwr[] (float64) + repeats_done, result == single-stream oracle."""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, N, R, out_path):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = importlib.import_module("rtl-power-fftw_b200")
    sh = importlib.import_module("rtl-power-fftw_b200.sharding")
    b0, b1 = sh.shard_byte_range(N, R, rank, world)
    first, last = sh.shard_range(R, rank, world)
    dev = torch.empty(b1 - b0, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        pkg.synth_fill_device(dev.data_ptr(), b1 - b0, b0, seed=4242, mode=1, device=rank, stream=s.cuda_stream)
        ctx = pkg.Context(N, device=rank)
        ctx.set_stream(s.cuda_stream)
        ctx.reset(last - first)
        ctx.push_device(dev.data_ptr(), b1 - b0)
        pwr = torch.zeros(N, dtype=torch.float64, device="cuda")
        done = ctx.finish_device(pwr.data_ptr())
        pwr, total = sh.allreduce_pwr(pwr, done)
        s.synchronize()
    if rank == 0:
        np.save(out_path, np.concatenate([[total], pwr.cpu().numpy()]))
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("N,R", [(4096, 4001), (65536, 37)])
def test_sharded_integration_nccl(tmp_path, N, R):
    import torch.multiprocessing as mp
    import oracle_lib as ol
    world = min(_ngpu(), 8)
    out = str(tmp_path / "res.npy")
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, N, R, out), nprocs=world, join=True)
    res = np.load(out)
    pkg = importlib.import_module("rtl-power-fftw_b200")
    whole = pkg.synth_fill_host(2 * N * R, 0, seed=4242, mode=1)
    po, done = ol.oracle_run_stream(N, R, whole, buf_length=whole.size, fft_kind=1 if N * R > 1 << 22 else 0)
    assert int(res[0]) == done == R
    assert np.max(np.abs(res[1:] - po) / po) < 1e-5
