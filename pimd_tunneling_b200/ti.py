"""Thermodynamic-integration bookkeeping around the propagation call: task layout and statistics of
pimd_par.f90:109-110, 243-257, 281-295, 379, 397-424.  Trajectories (lambda point x repetition) are
independent units; across GPUs they are block-partitioned by global id and the only exchange is ONE
all-reduce of {sum I, sum I^2, count} per lambda point (replacing MPI_Gather, pimd_par.f90:389)."""
import ctypes

import numpy as np

from ._lib import check, hptr, lib


def global_ids(nintegral, nrep):
    """id = nrep*(ilambda-1) + irep (0-based), pimd_par.f90:249-253"""
    return np.arange(nintegral * nrep, dtype=np.int64)


def shard(ntotal, rank, world):
    """Block rule of pimd_par.f90:109-110,281-295: rank r owns [r*ncalcs, (r+1)*ncalcs) with
    ncalcs = ceil(ntotal/world); padded tasks of the reference are simply absent here."""
    ncalcs = -(-ntotal // world)
    lo = min(rank * ncalcs, ntotal)
    hi = min(lo + ncalcs, ntotal)
    return lo, hi


def partial_sums(dHdr, traj_gid, nrep, nintegral, betan):
    dHdr = np.ascontiguousarray(dHdr, dtype=np.float64)
    gid = np.ascontiguousarray(traj_gid, dtype=np.int64)
    sums = np.zeros((nintegral, 3))
    check(lib().pimdk_ti_partial_sums(dHdr.size, hptr(dHdr), hptr(gid), nrep, nintegral, float(betan), hptr(sums)))
    return sums


def comm_info():
    """(rank, nranks, nccl_version) of the library's own communicator; (0, 1, 0) before comm_init"""
    r, n, v = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
    check(lib().pimdk_comm_info(ctypes.byref(r), ctypes.byref(n), ctypes.byref(v)))
    return int(r.value), int(n.value), int(v.value)


def comm_init(rank, world, exchange=None):
    """Create the library's NCCL communicator (pimdk_comm_init): rank 0 makes the 128-byte unique id and `exchange`
    hands it to the other ranks — a callable id_bytes_or_None -> id_bytes.  Default: a torch.distributed broadcast
    over whatever process group the launcher set up (plumbing only; the data-path collective is the library's)."""
    if world == 1:
        check(lib().pimdk_comm_init(0, 1, None))
        return
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        check(lib().pimdk_comm_unique_id(buf))
    if exchange is None:
        import torch
        import torch.distributed as dist

        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
    else:
        raw = exchange(buf.raw if rank == 0 else None)
    check(lib().pimdk_comm_init(rank, world, ctypes.create_string_buffer(raw, 128)))


def comm_finalize():
    check(lib().pimdk_comm_finalize())


def reduce_dev(ntraj, dHdr_ptr, gid_ptr, nrep, nintegral, betan):
    """Per-lambda {sum I, sum I^2, count} of the device-resident dHdr of the last propagate_dev call, summed over the
    ranks of the library's communicator (one ncclAllReduce of 3*nintegral doubles); returns sums(nintegral, 3)."""
    sums = np.zeros((nintegral, 3))
    check(lib().pimdk_ti_reduce_dev(ntraj, dHdr_ptr, gid_ptr, nrep, nintegral, float(betan), hptr(sums)))
    return sums


def allreduce_sums(sums):
    """The one collective of the path.  With the library's communicator initialised (comm_init): ncclAllReduce inside
    libpimdk.so.  Otherwise, if a torch.distributed group exists (the world-size-2 gloo rig of the CPU tests: no GPU, hence
    no NCCL), the reduction alone falls back to it.  One rank: nothing to do."""
    if comm_info()[1] > 1:
        out = np.ascontiguousarray(sums, dtype=np.float64).copy()
        check(lib().pimdk_ti_allreduce(out.shape[0], hptr(out)))
        return out
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sums
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.ascontiguousarray(sums)).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def finish(sums, weights, betan):
    """mean/variance per lambda, Delta A, sigma_A, q/q0 (pimd_par.f90:401-424)"""
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    n = w.size
    mean = np.empty(n)
    var = np.empty(n)
    dA = ctypes.c_double()
    sA = ctypes.c_double()
    q = ctypes.c_double()
    check(lib().pimdk_ti_finish(n, hptr(sums), hptr(w), float(betan), hptr(mean), hptr(var), ctypes.addressof(dA),
                                ctypes.addressof(sA), ctypes.addressof(q)))
    return {"mean": mean, "var": var, "deltaA": dA.value, "sigmaA": sA.value, "q_over_q0": q.value}
