"""Communicator: the role mpi4py's COMM_WORLD plays in the reference
(SURVEY.md 2b), on top of torch.distributed -- NCCL over NVLink between the
B200s of one box, gloo for the CPU tests of the host logic.

One process per GPU.  Only the operations the hot path needs exist:
  * `exchange`  : grouped point-to-point sends/receives of device buffers
                  (halo slices, wavelet boundary slices;
                  mpi_vector.py:140-203's Isend/Irecv),
  * `all_to_all`: the time<->space transpose (mpi_vector.py:212-240),
  * `allreduce_sum`: Krylov scalars (mpi_vector.py:209),
  * object bcast/gather/Barrier for setup and statistics.
"""
import time

import torch


def Wtime():
    return time.perf_counter()


# The reference's `time_applies` / `time_communication` are wall times of
# blocking NumPy / MPI calls.  Kernels and NCCL calls are asynchronous here, so
# the drivers that REPORT those fields (heateq_mpi.main, heateq_mpi_timing)
# switch this on: every timed bracket then drains the device on both sides and
# the fields keep the reference's meaning (time until the result exists).  It
# stays off in library use and in bench.py (no host synchronisation inside the
# Krylov loop).
SYNC_TIMING = False


def Wtime_device():
    """Wall clock for a timed bracket around device work."""
    if SYNC_TIMING and torch.cuda.is_available():
        torch.cuda.synchronize()
    return time.perf_counter()


class SerialComm:
    """One rank, no library underneath (python3 heateq.py-style runs)."""
    rank, size = 0, 1

    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Split_type(self, *_):
        return self

    def allreduce_sum(self, t):
        return t

    def allreduce(self, obj):
        return obj

    def bcast(self, obj, root=0):
        return obj

    def gather(self, obj, root=0):
        return [obj]

    def Barrier(self):
        pass

    def exchange(self, sends, recvs):
        # a rank never messages itself in the plans built by timeop.py
        assert not sends and not recvs

    def exchange_begin(self, sends, recvs):
        assert not sends and not recvs
        return []

    def exchange_end(self, reqs):
        pass

    def all_to_all(self, send_chunks, recv_chunks):
        recv_chunks[0].copy_(send_chunks[0])


class TorchComm:
    """torch.distributed process group (nccl on GPUs, gloo on CPU)."""
    def __init__(self, group=None):
        import torch.distributed as dist
        assert dist.is_initialized()
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)
        self._peer = None

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def Split_type(self, *_):
        return self  # one box: every rank shares the node

    def allreduce_sum(self, t):
        """In-place sum of a tensor over the ranks; returns it."""
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce(self, obj):
        """Python-number sum (the reference's lowercase comm.allreduce)."""
        objs = [None] * self.size
        self.dist.all_gather_object(objs, obj, group=self.group)
        total = objs[0]
        for o in objs[1:]:
            total = total + o
        return total

    def bcast(self, obj, root=0):
        box = [obj]
        self.dist.broadcast_object_list(box, src=root, group=self.group)
        return box[0]

    def gather(self, obj, root=0):
        objs = [None] * self.size
        self.dist.all_gather_object(objs, obj, group=self.group)
        return objs if self.rank == root else None

    def Barrier(self):
        self.dist.barrier(group=self.group)

    def exchange(self, sends, recvs):
        """sends / recvs: {peer: contiguous tensor}.  All transfers are posted
        as one batch (a single NCCL group) and waited for."""
        self.exchange_end(self.exchange_begin(sends, recvs))

    def exchange_begin(self, sends, recvs):
        """Posts the batch and returns its requests without waiting: with NCCL
        the transfers run on the communicator's own stream, ordered after the
        work already enqueued on the current stream; what the caller enqueues
        before `exchange_end` overlaps them (the `callback` of
        mpi_vector.py:155-183)."""
        if peer_copy_enabled() and all(t.is_cuda for t in list(sends.values()) +
                                       list(recvs.values())):
            if self._peer is None:
                self._peer = PeerWindows(self)
            return self._peer.exchange_begin(sends, recvs)
        ops = []
        for peer in sorted(recvs):
            ops.append(self.dist.P2POp(self.dist.irecv, recvs[peer], peer,
                                       self.group))
        for peer in sorted(sends):
            ops.append(self.dist.P2POp(self.dist.isend, sends[peer], peer,
                                       self.group))
        self.bytes_sent = getattr(self, 'bytes_sent', 0) + sum(
            t.numel() * t.element_size() for t in sends.values())
        return self.dist.batch_isend_irecv(ops) if ops else []

    def exchange_end(self, reqs):
        """The current stream (NCCL) / the host (gloo) waits for the batch."""
        for req in reqs:
            req.wait()


    def all_to_all(self, send_chunks, recv_chunks):
        """send_chunks[p] goes to rank p, recv_chunks[p] comes from rank p."""
        if send_chunks[0].is_cuda:
            self.dist.all_to_all(recv_chunks, send_chunks, group=self.group)
        else:  # gloo has no all_to_all: post the pairwise transfers
            me = self.rank
            recv_chunks[me].copy_(send_chunks[me])
            self.exchange(
                {p: send_chunks[p] for p in range(self.size) if p != me},
                {p: recv_chunks[p] for p in range(self.size) if p != me})



def peer_copy_enabled():
    import os
    return os.environ.get('STK_PEER_COPY', '0') == '1'


class _StreamEvent:
    """`wait()` makes the current stream wait for an event of the transfer
    stream (the request protocol of exchange_end)."""
    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class PeerWindows:
    """Point-to-point batches as peer-to-peer copies over NVLink: every rank
    owns two send windows in HBM that the other ranks of the box map through
    CUDA IPC; an exchange packs into the window, one scalar allreduce tells
    everyone that all windows are written, and each rank pulls its pieces from
    its peers' windows with device-to-device copies (copy engines, no SMs, no
    staging through NCCL's channel buffers).  The windows alternate, so the
    allreduce of the next exchange also guards the reuse.  Everything runs on a
    side stream; `exchange_end` makes the caller's stream wait for it.

    Needs all ranks on one node with every GPU visible to every process (what
    torchrun gives).  The pieces' offsets inside the senders' windows are
    agreed on once per exchange pattern (an object all-gather)."""
    def __init__(self, comm):
        self.comm = comm
        self.stream = torch.cuda.Stream()
        self.capacity = 0
        self.windows = None      # my two windows (uint8)
        self.peer = None         # peer[p][k]: rank p's window k as seen from here
        self.layouts = {}
        self.count = 0
        self.flag = torch.zeros(1, dtype=torch.float32, device='cuda')

    def _ensure(self, nbytes):
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        need = torch.tensor([nbytes], dtype=torch.int64, device='cuda')
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.comm.group)
        need = int(need.item())
        if need <= self.capacity:
            return
        torch.cuda.synchronize()
        self.peer = None
        self.capacity = int(need * 1.25) + 4096
        self.windows = [torch.empty(self.capacity, dtype=torch.uint8, device='cuda')
                        for _ in range(2)]
        mine = [reduce_tensor(w) for w in self.windows]
        everyone = [None] * self.comm.size
        dist.all_gather_object(everyone, mine, group=self.comm.group)
        self.peer = []
        for p, handles in enumerate(everyone):
            if p == self.comm.rank:
                self.peer.append(self.windows)
            else:
                self.peer.append([fn(*args) for fn, args in handles])
        self.count = 0

    def exchange_begin(self, sends, recvs):
        import torch.distributed as dist
        key = (tuple(sorted((p, t.numel() * t.element_size()) for p, t in sends.items())),
               tuple(sorted((p, t.numel() * t.element_size()) for p, t in recvs.items())))
        if key not in self.layouts:
            # offsets of my pieces in my window, 256-byte aligned; every rank
            # learns where its pieces sit in its peers' windows
            off, mine = 0, {}
            for p in sorted(sends):
                n = sends[p].numel() * sends[p].element_size()
                mine[p] = (off, n)
                off += (n + 255) // 256 * 256
            table = [None] * self.comm.size
            dist.all_gather_object(table, mine, group=self.comm.group)
            self._ensure(off)
            src = {}
            for p in recvs:
                o, n = table[p][self.comm.rank]
                assert n == recvs[p].numel() * recvs[p].element_size()
                src[p] = o
            self.layouts[key] = (mine, src)
        mine, src = self.layouts[key]
        k = self.count & 1
        self.count += 1
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            win = self.windows[k]
            for p, t in sends.items():
                o, n = mine[p]
                win[o:o + n].view(t.dtype).view(t.shape).copy_(t)
            # all windows written (and the previous use of the other window read)
            dist.all_reduce(self.flag, op=dist.ReduceOp.MAX, group=self.comm.group)
            for p, t in recvs.items():
                n = t.numel() * t.element_size()
                t.copy_(self.peer[p][k][src[p]:src[p] + n].view(t.dtype).view(t.shape),
                        non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        for t in list(sends.values()) + list(recvs.values()):
            t.record_stream(self.stream)
        self.comm.bytes_sent = getattr(self.comm, 'bytes_sent', 0) + sum(
            t.numel() * t.element_size() for t in sends.values())
        return [_StreamEvent(done)]


_world = None


def world():
    """COMM_WORLD: the default torch.distributed group when initialised (one
    process per GPU under torchrun), else a serial communicator."""
    global _world
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        if not isinstance(_world, TorchComm):
            _world = TorchComm()
    elif _world is None or isinstance(_world, TorchComm):
        _world = SerialComm()
    return _world


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK,
    LOCAL_RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT) and bind this process to
    its GPU.  No-op for single-process runs."""
    import os
    import torch.distributed as dist
    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # The exchanges of this path are point-to-point batches between a few
        # peers (halo slices to the two neighbours, wavelet boundary slices to
        # <= 7 ranks), 8-130 MB per peer.  NCCL's default gives a peer pair few
        # channels (measured 60-160 GB/s of the 770 GB/s a B200 pair can move);
        # more channels per peer let one pair use the NVLink bandwidth.
        os.environ.setdefault('NCCL_MIN_P2P_NCHANNELS', '16')
        dist.init_process_group(backend=backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    return world()
