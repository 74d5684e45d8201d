"""CPU tests of the host-side rendezvous of the multi-GPU paths (csrc/comm.cpp, spmv_b200_comm_*):
one PROCESS per rank over an abstract unix-domain socket -- all-gather, barrier, and file descriptors
passed as SCM_RIGHTS (what carries the cuMem allocations of the symmetric rank vectors between the
processes on the GPU box).  World size 3; no GPU involved."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import ctypes as C, os, struct, sys
    sys.path.insert(0, %(root)r)
    from _load_pkg import load_pkg
    sp = load_pkg()
    rank, world, session = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    comm = C.c_void_p()
    assert sp.lib.spmv_b200_comm_create(rank, world, session.encode(), 30, C.byref(comm)) == 0
    # all-gather of a 12-byte record
    mine = struct.pack("<iq", rank, 1000 + rank * rank)
    recv = C.create_string_buffer(12 * world)
    assert sp.lib.spmv_b200_comm_allgather(comm, mine, recv, 12) == 0
    got = [struct.unpack_from("<iq", recv.raw, 12 * p) for p in range(world)]
    assert got == [(p, 1000 + p * p) for p in range(world)], got
    for _ in range(3):
        assert sp.lib.spmv_b200_comm_barrier(comm) == 0
    # descriptor passing: every rank offers the read end of a pipe it has written its rank into
    r, w = os.pipe()
    os.write(w, b"rank%%d" %% rank)
    os.close(w)
    fds = (C.c_int * world)()
    assert sp.lib.spmv_b200_comm_allgather_fds(comm, r, fds) == 0
    os.close(r)
    assert sp.lib.spmv_b200_comm_barrier(comm) == 0
    # each pipe holds one message and world readers: rank p reads the pipe of rank (p + 1) %% world
    target = (rank + 1) %% world
    assert len(set(fds)) == world and all(fd >= 0 for fd in fds)
    assert os.read(fds[target], 16) == b"rank%%d" %% target
    for fd in fds:
        os.close(fd)
    assert sp.lib.spmv_b200_comm_barrier(comm) == 0
    sp.lib.spmv_b200_comm_destroy(comm)
    print("ok", rank)
""") % {"root": ROOT}


def test_socket_comm_three_processes(sp, tmp_path):
    world = 3
    session = f"pytest-{os.getpid()}"
    procs = [subprocess.Popen([sys.executable, "-c", WORKER, str(r), str(world), session], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=120) for p in procs]
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"ok {r}" in out, (r, out, err[-2000:])


def test_single_rank_comm_is_trivial(sp):
    import ctypes as C
    comm = C.c_void_p()
    assert sp.lib.spmv_b200_comm_create(0, 1, b"solo", 5, C.byref(comm)) == 0
    recv = C.create_string_buffer(4)
    assert sp.lib.spmv_b200_comm_allgather(comm, b"abcd", recv, 4) == 0 and recv.raw == b"abcd"
    assert sp.lib.spmv_b200_comm_barrier(comm) == 0
    sp.lib.spmv_b200_comm_destroy(comm)
    assert sp.lib.spmv_b200_comm_create(2, 2, b"bad", 5, C.byref(comm)) != 0  # rank out of range
