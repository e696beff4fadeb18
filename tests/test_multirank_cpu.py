"""World-size-2 gloo test of the N>1 host logic (sharding by spectrum index +
one allreduce of pwr[] and repeats_done).  Per-rank spectra come from the CPU
oracle here -- on the GPU box the same sharding code drives the CUDA path
(tests/test_gpu_multirank.py, bench.py --gpus N)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, N, R, out_path):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("rtl-power-fftw_b200")
    sh = importlib.import_module("rtl-power-fftw_b200.sharding")
    import oracle_lib as ol2
    b0, b1 = sh.shard_byte_range(N, R, rank, world)
    # each rank regenerates only its slice of the global stream (counter-based source)
    mine = pkg.synth_fill_host(b1 - b0, b0, seed=99, mode=1)
    first, last = sh.shard_range(R, rank, world)
    pwr, done = ol2.oracle_run_stream(N, last - first, mine, buf_length=max(16384, mine.size))
    t = torch.from_numpy(pwr.copy())
    t, total_done = sh.allreduce_pwr(t, done)
    if rank == 0:
        np.save(out_path, np.concatenate([[total_done], t.numpy()]))
    dist.destroy_process_group()


def test_two_rank_sharded_integration_equals_single(tmp_path):
    N, R, world = 512, 37, 2          # odd spectrum count: uneven shards
    out = str(tmp_path / "res.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, N, R, out), nprocs=world, join=True)
    res = np.load(out)
    pkg = importlib.import_module("rtl-power-fftw_b200")
    whole = pkg.synth_fill_host(2 * N * R, 0, seed=99, mode=1)
    pwr, done = ol.oracle_run_stream(N, R, whole, buf_length=whole.size)
    assert int(res[0]) == done == R
    assert np.max(np.abs(res[1:] - pwr) / pwr) < 1e-12
