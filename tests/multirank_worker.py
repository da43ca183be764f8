"""Worker of tests/test_gpu_multirank.py (launched under torchrun, one process per GPU, NCCL): the marker-sharded step
and the sharded forward search must give, bit for bit, what one GPU gives on the whole data set."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from eagleeverything_b200 import am, device, synth  # noqa: E402
from eagleeverything_b200 import dist as egd  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device.init(local)
    n, L = int(sys.argv[1]), int(sys.argv[2])
    G = synth.genotypes(n, L)                                    # every rank generates the same data set
    S, V, a_hat = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in synth.scan_inputs(n))

    def stores(c0, c1):
        img = np.concatenate([synth.ascii_image(G[:, c0:c1]).reshape(-1), np.zeros(64, np.uint8)])
        kb, err = device.decode_kb(torch.from_numpy(img).cuda(), c1 - c0 + 1, n, c1 - c0)
        assert int(err[0].item()) == 0
        return kb, device.transpose_kb(kb, n, c1 - c0)

    # ---- the whole data set on this GPU alone
    kb1, t1 = stores(0, L)
    K1 = device.mmt_finalize(device.syrk_kb(kb1, n, L), n)
    Wp1 = device.scan_prepare(S, V, a_hat, n)
    a1, v1 = device.scan(t1, L, n, Wp1)
    b1, i1 = device.argmax_tsq(a1, v1)
    # ---- the sharded step
    c0, c1 = egd.shard_range(L, world, rank)
    kb, t = stores(c0, c1)
    C32 = device.syrk_kb(kb, n, c1 - c0)
    egd.allreduce_partial_mmt(C32)
    K = device.mmt_finalize(C32, n)
    assert torch.equal(K, K1), "all-reduced K differs from the single-GPU K"
    Wp = device.scan_prepare_sharded(S, V, a_hat, n, rank, world)
    assert torch.equal(Wp, Wp1), "sharded pre-products differ from the single-GPU ones"
    a, v = device.scan(t, c1 - c0, n, Wp)
    assert torch.equal(a, a1[c0:c1]) and torch.equal(v, v1[c0:c1]), "a / var(a) of the shard differ"
    best, idx = device.argmax_tsq(a, v)
    assert egd.global_argmax(best, idx, c0)[1] == int(i1.item()), "pick differs"
    a_all = egd.gather_sharded(a, L, world)
    assert torch.equal(a_all, a1)
    # ---- the sharded forward search against the single-GPU one
    y, _ = synth.phenotype(G)
    r1 = am.AM_resident(kb1, t1, n, L, y, maxit=5)
    rs = am.AM_resident(kb, t, n, L, y, maxit=5, shard=egd.Shard(L, world, rank))
    assert rs["all_picked"] == r1["all_picked"] and rs["selected"] == r1["selected"], (rs["all_picked"], r1["all_picked"])
    np.testing.assert_allclose(rs["extBIC"], r1["extBIC"], rtol=1e-12)
    dist.barrier()
    if rank == 0:
        print("MULTIRANK_OK", world, r1["all_picked"])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
