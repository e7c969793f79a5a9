"""N>1 host logic of the Monte-Carlo driver on CPU: two gloo ranks, fake per-rank launch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    for p in (PKG, REPO):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mc_driver import run_intervals, split_frames

    seen = []

    def fake_launch(frames_local, counters, frame_offset):
        # deterministic stand-in for ldpc_mc_run: frame g fails iff g % 10 == 0 (3 bit errors each)
        seen.append((frames_local, frame_offset))
        for g in range(frame_offset, frame_offset + frames_local):
            counters[0] += 1
            if g % 10 == 0:
                counters[1] += 1
                counters[2] += 3
            else:
                counters[3] += g % 4
                counters[4] += 1

    # exact frame budget, several intervals
    total, cursor = run_intervals(fake_launch, device="cpu", rank=rank, world=world, group=None, distributed=True,
                                  frames=1001, interval_frames=250)
    assert cursor == 1001
    assert total.frames == 1001 and total.frame_errors == 101 and total.bit_errors == 303
    assert total.conv_count == 900 and total.conv_sum == sum(g % 4 for g in range(1001) if g % 10)
    # every interval was split contiguously and this rank saw only its share
    offs = 0
    for (fl, fo), chunk in zip(seen, (250, 250, 250, 250, 1)):
        lo, hi = split_frames(chunk, rank, world)
        assert (fl, fo) == (hi - lo, offs + lo)
        offs += chunk
    # stopping rule on the REDUCED counters: both ranks stop after the same interval
    seen.clear()
    total2, cursor2 = run_intervals(fake_launch, device="cpu", rank=rank, world=world, group=None, distributed=True,
                                    max_frames=100000, min_frame_errors=40, interval_frames=100, frame_cursor=0)
    assert total2.frame_errors == 40 and total2.frames == 400 and cursor2 == 400 and len(seen) == 4
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
        f.write("ok")
    dist.destroy_process_group()


def test_two_rank_counter_allreduce_and_stop_rule(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_single_rank_needs_no_process_group():
    from mc_driver import run_intervals

    def launch(frames_local, counters, frame_offset):
        counters[0] += frames_local

    total, cursor = run_intervals(launch, device="cpu", rank=0, world=1, group=None, distributed=False, frames=37)
    assert total.frames == 37 and cursor == 37
    with pytest.raises(ValueError):
        run_intervals(launch, device="cpu", rank=0, world=1, group=None, distributed=False)
