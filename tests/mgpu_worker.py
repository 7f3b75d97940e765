"""Worker of tests/test_multi_gpu.py (one process per GPU under torch.distributed.run): the library's own NCCL communicator
(pimdk_comm_init) all-reduces the per-lambda estimator sums of a sharded TI batch; rank 0 also runs the whole batch alone and
checks that the sharded job reproduces its statistics (results depend on global trajectory ids only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pimd_tunneling_b200 as pk  # noqa: E402
from bench import ti_path, wells  # noqa: E402
from pimd_tunneling_b200 import path as P  # noqa: E402
from pimd_tunneling_b200 import ti  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")            # plumbing only: hands the unique id over; the collective under test is the library's
pk.init(local)
ti.comm_init(rank, world)
assert ti.comm_info()[:2] == (rank, world) and ti.comm_info()[2] > 20000
pes = pk.McmodMass("2dtest").V_init()
a, b, mass = wells("2dtest")
n, nintegral, nrep, steps = 160, 4, 6, 12
vi = pk.VerletInt(pes, n, mass, 10.0, dt=1e-3, NMC=steps, Noutput=5, seed=77).init_nm()
lam, path, spl = ti_path("2dtest", a, b)
xi, w = vi.gauleg(0.0, 1.0, nintegral)
xint, dbdxi = P.endpoints(lam, path, spl, xi)


def run(gid):
    il = gid // nrep
    x, p = vi.init_path(xi[il], lam, path, spl, traj_gid=gid)
    _, _, dH = vi.propagate_pimd_nm(x, p, a, np.asfortranarray(xint[:, :, il]), np.asfortranarray(dbdxi[:, :, il]), traj_gid=gid)
    return dH


ids = ti.global_ids(nintegral, nrep)
lo, hi = ti.shard(ids.size, rank, world)
dH = run(ids[lo:hi])
# (1) host sums -> pimdk_ti_allreduce
sums = ti.allreduce_sums(ti.partial_sums(dH, ids[lo:hi], nrep, nintegral, vi.betan))
# (2) device sums -> pimdk_ti_reduce_dev
d = torch.from_numpy(dH).cuda()
g = torch.from_numpy(ids[lo:hi].copy()).cuda()
sums_dev = ti.reduce_dev(hi - lo, d.data_ptr(), g.data_ptr(), nrep, nintegral, vi.betan)
if rank == 0:
    ref = ti.partial_sums(run(ids), ids, nrep, nintegral, vi.betan)
    assert np.array_equal(sums[:, 2], ref[:, 2]) and np.abs(sums - ref).max() <= 1e-12 * np.abs(ref).max(), (sums, ref)
    assert np.abs(sums_dev - ref).max() <= 1e-12 * np.abs(ref).max()
    out = ti.finish(sums, w, vi.betan)
    assert np.isfinite(out["deltaA"])
    print("mgpu ok: %d ranks, NCCL %d, deltaA %.6f" % (world, ti.comm_info()[2], out["deltaA"]))
ti.comm_finalize()
dist.destroy_process_group()
pk.finalize()
