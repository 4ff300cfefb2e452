"""CPU tier, world size 2 over gloo: the N > 1 host logic of the column-sharded msa2eds path — shard plan,
halo windows, the all-gather of byte counts, offset writes into one file pair. The 'device' is the
test-only emulator build of the kernels; on GPUs the same code runs over NCCL (bench.py --gpus N)."""
import os
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R, C, WRAP, L, SEED, PPM, HALO = 6, 1500, 60, 10, 3, 40000, 64


def _rank_main(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    import emu_lib
    from edsparser_b200 import shard

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = emu_lib.lib().context()
        ctx.set_tuning(2, 1)
        lo, hi, wb, we = shard.plan(C, world, rank, HALO)
        view = ctx.msa_synth(R, C, WRAP, col_begin=wb, col_count=we - wb, seed=SEED, variable_ppm=PPM)
        view.own_begin, view.own_end = lo, hi
        e, s, st = ctx.msa_transform_device(view, L)
        eds, seds = ctx.download(e), ctx.download(s)
        eo, so, et, stot = shard.gather_offsets(dist, torch.device("cpu"), len(eds), len(seds))
        ex = shard.OffsetExchange(dist, torch.device("cpu"))  # the queued form used by bench.py gives the same offsets
        ex.post(len(eds), len(seds))
        assert ex.offsets() == (eo, so, et, stot)
        dist.barrier()
        shard.write_slice(os.path.join(outdir, "out.leds"), eo, eds, et, rank)
        shard.write_slice(os.path.join(outdir, "out.seds"), so, seds, stot, rank)
        dist.barrier()
        ctx.close()
    finally:
        dist.destroy_process_group()


def test_two_ranks_write_one_file_pair():
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_lib
    import oracle_lib
    from edsparser_b200 import synth

    emu_lib.lib()  # build before the ranks race for it
    with tempfile.TemporaryDirectory() as outdir:
        port = 29000 + os.getpid() % 2000
        mp.spawn(_rank_main, args=(2, port, outdir), nprocs=2, join=True)
        text = synth.fasta_window(R, C, WRAP, seed=SEED, variable_ppm=PPM)
        exp = oracle_lib.msa2eds(text, L)
        with open(os.path.join(outdir, "out.leds"), "rb") as f:
            assert f.read() == exp[0]
        with open(os.path.join(outdir, "out.seds"), "rb") as f:
            assert f.read() == exp[1]


def test_plan_covers_every_column_once():
    from edsparser_b200 import shard

    for total, world in ((10, 1), (1000, 3), (7, 8), (30_000_000, 8)):
        cuts = [shard.plan(total, world, r, 5) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == total
        for a, b in zip(cuts[:-1], cuts[1:]):
            assert a[1] == b[0]
        for lo, hi, wb, we in cuts:
            assert 0 <= wb <= lo <= hi <= we <= total
