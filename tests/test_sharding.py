"""Host-side multi-GPU logic, exercised on CPU with a world_size-2 gloo group (the data path has no
collective: ranks own disjoint contiguous row ranges; only the timing is reduced)."""
import os
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_partition_covers_rows(wf):
    from watfft_b200.sharding import partition
    for batch in (0, 1, 7, 8, 262144, 262145):
        for world in (1, 2, 3, 4, 8):
            b = partition(batch, world)
            assert b[0][0] == 0 and b[-1][1] == batch
            assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
            assert partition(batch, world, world - 1) == b[-1]
    # config 5: 262144 transforms of N=4096 over 2/4/8 GPUs
    assert [e - s for s, e in partition(262144, 8)] == [32768] * 8
    with pytest.raises(ValueError):
        partition(8, 0)


def test_rank_device_orders(wf):
    from watfft_b200.sharding import rank_device
    assert [rank_device(i, 2, 8) for i in range(2)] == [0, 4]            # one rank per host bridge
    assert [rank_device(i, 4, 8) for i in range(4)] == [0, 2, 4, 6]
    assert [rank_device(i, 8, 8) for i in range(8)] == list(range(8))
    assert [rank_device(i, 2, 2) for i in range(2)] == [0, 1]            # CUDA_VISIBLE_DEVICES already narrowed the box
    assert [rank_device(i, 3, 8) for i in range(3)] == [0, 1, 2]         # no even spread: sequential
    assert [rank_device(i, 2, 8, "seq") for i in range(2)] == [0, 1]
    assert rank_device(0, 1, 8) == 0
    for w in (2, 4, 8):                                                  # distinct GPUs whatever the order
        assert len({rank_device(i, w, 8) for i in range(w)}) == w
    with pytest.raises(ValueError):
        rank_device(2, 2, 8)


def test_grains_tile_the_batch(wf):
    from watfft_b200.sharding import grains
    for batch in (0, 1, 255, 256, 1000, 10007, 262144):
        for row_bytes in (64, 1024, 16384):
            for grain_bytes in (64 << 10, 1 << 20, 32 << 20):
                g, lst = grains(batch, row_bytes, grain_bytes)
                assert g >= 1 and (g < 512 or g % 256 == 0)
                assert sum(r for _, r in lst) == batch
                assert all(r0 == i * g for i, (r0, _) in enumerate(lst))
                assert all(r == g for _, r in lst[:-1]) and (not lst or 1 <= lst[-1][1] <= g)
    assert grains(262144, 16384, 32 << 20) == (2048, [(i * 2048, 2048) for i in range(128)])     # configs[4]: 128 grains
    with pytest.raises(ValueError):
        grains(10, 0)


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {str(ROOT)!r})
        import torch, torch.distributed as dist
        import numpy as np
        import watfft_b200
        from watfft_b200.sharding import partition, max_over_ranks
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        batch = 1001
        b, e = partition(batch, world, rank)
        # every rank reports the rows it owns; together they must tile [0, batch) exactly once
        owned = torch.zeros(batch, dtype=torch.int32)
        owned[b:e] = 1
        dist.all_reduce(owned)
        assert int(owned.min()) == 1 and int(owned.max()) == 1
        # job time = max over ranks
        t = max_over_ranks(1.0 + rank)
        assert t == float(world)
        # whole-job throughput the way bench.py forms it
        rows = torch.tensor([e - b], dtype=torch.int64)
        dist.all_reduce(rows)
        assert int(rows.item()) == batch
        dist.destroy_process_group()
        sys.stdout.write("rank %d ok\\n" % rank)      # one write: the two ranks share the pipe
        sys.stdout.flush()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    import socket
    with socket.socket() as sk:                       # a free rendezvous port (a fixed one collides with leftovers)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", ["static", "dynamic"])
def test_sharded_split_fft_one_process(wf, oracle, schedule):
    """ShardedSplitFFT from one process over every visible device (a 1-GPU box shards over GPU 0 twice): every row is
    transformed exactly once whichever device or worker took it, including the short last grain."""
    import numpy as np
    import torch
    from conftest import f32_bound, rel_err
    from watfft_b200.sharding import ShardedSplitFFT
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0]
    n, batch = 256, 10007                                   # 1 KiB rows; 64 KiB grains -> 157 grains, the last one 23 rows
    rng = np.random.default_rng(3)
    re = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    im = rng.uniform(-1, 1, (batch, n)).astype(np.float32)
    sh = ShardedSplitFFT(n, batch, devices, schedule=schedule, grain_bytes=64 << 10, workers_per_device=2)
    sh.scatter(re, im)
    sh.run()
    g_re, g_im = sh.gather()
    want = np.fft.fft(re.astype(np.float64) + 1j * im.astype(np.float64), axis=1)
    assert np.max(np.abs((g_re + 1j * g_im) - want)) / np.sqrt(2 * n) <= f32_bound(n)
    for r in (0, batch // 2, batch - 1):
        o_re, o_im = oracle.fft_split_f32(re[r], im[r])
        assert rel_err(np.r_[g_re[r], g_im[r]], np.r_[o_re, o_im], np.r_[re[r], im[r]]) <= f32_bound(n)
    if schedule == "dynamic":
        assert sum(sh.last_counts) == len(sh.grains) == -(-batch // sh.grain)
    sh.run(inverse=True)
    b_re, b_im = sh.gather()
    assert np.max(np.abs(b_re - re)) < 1e-5 and np.max(np.abs(b_im - im)) < 1e-5
    sh.dispose()
