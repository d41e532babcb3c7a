"""TEST INFRASTRUCTURE ONLY -- stand-in for `mpi4py.MPI` (see package docstring).

Message matching follows MPI rules closely enough for the reference:
messages between a (source, dest) pair with equal tags are non-overtaking.
Sends are buffered (copied at Isend time), receives complete in Waitall/Recv.
"""
import time

import numpy as np

DOUBLE = 'MPI_DOUBLE'
COMM_TYPE_SHARED = 'MPI_COMM_TYPE_SHARED'
ANY_TAG = -1


def Wtime():
    return time.perf_counter()


class _DoneRequest:
    def Wait(self):
        return None


class _RecvRequest:
    def __init__(self, comm, buf, source, tag):
        self.comm, self.buf, self.source, self.tag = comm, buf, source, tag

    def Wait(self):
        self.comm._complete_recv(self.buf, self.source, self.tag)


class Request:
    @staticmethod
    def Waitall(reqs):
        # Receives are completed in posting order; sends are already buffered.
        for r in reqs:
            r.Wait()


class _Comm:
    """World communicator.  `_links` is None for the 1-rank in-process world,
    otherwise {peer_rank: multiprocessing.Connection}."""
    def __init__(self, rank=0, size=1, links=None):
        self.rank, self.size = rank, size
        self._links = links
        self._self_queue = []  # messages a rank sends to itself
        self._stash = {p: [] for p in range(size)}  # received, unmatched
        self._outbox = {}  # peer -> queue drained by a sender thread

    # -- introspection
    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def Split_type(self, split_type):
        return self  # every emulated rank lives on the same "node"

    # -- point to point
    def _post(self, dest, tag, payload):
        dest = int(dest)
        if dest == self.rank:
            self._self_queue.append((tag, payload))
        else:
            self._send(dest, (tag, payload))

    def _send(self, dest, msg):
        """Non-blocking send: pipes have a small kernel buffer, so two ranks
        that both send a large message before receiving would deadlock."""
        import queue
        import threading
        if dest not in self._outbox:
            q = queue.Queue()
            link = self._links[dest]

            def pump():
                while True:
                    link.send(q.get())
                    q.task_done()

            threading.Thread(target=pump, daemon=True).start()
            self._outbox[dest] = q
        self._outbox[dest].put(msg)

    def _pull(self, source, pred):
        """Next message from `source` satisfying pred (others are stashed)."""
        stash = self._stash[source]
        for k, msg in enumerate(stash):
            if pred(msg[0]):
                return stash.pop(k)
        while True:
            msg = self._links[source].recv()
            if pred(msg[0]):
                return msg
            stash.append(msg)

    def Isend(self, buf, dest, tag=0):
        self._post(dest, int(tag), np.array(buf, dtype=np.float64, copy=True))
        return _DoneRequest()

    def Irecv(self, buf, source, tag=ANY_TAG):
        return _RecvRequest(self, buf, int(source), int(tag))

    def Recv(self, buf, source, tag=ANY_TAG):
        self._complete_recv(buf, int(source), int(tag))

    def _complete_recv(self, buf, source, tag):
        def pred(t):
            return isinstance(t, int) and (tag == ANY_TAG or t == tag)

        if source == self.rank:
            payload = None
            for k, (t, pl) in enumerate(self._self_queue):
                if pred(t):
                    payload = self._self_queue.pop(k)[1]
                    break
            assert payload is not None, "self-recv without matching self-send"
        else:
            payload = self._pull(source, pred)[1]
        np.asarray(buf)[...] = payload.reshape(np.asarray(buf).shape)

    # -- object collectives (rank 0 is the hub)
    def _xchg_obj(self, obj, tag):
        """all-gather of python objects through rank 0."""
        if self.size == 1:
            return [obj]
        if self.rank == 0:
            objs = [obj]
            for p in range(1, self.size):
                objs.append(self._pull(p, lambda t: t == tag)[1])
            for p in range(1, self.size):
                self._send(p, (tag, objs))
            return objs
        self._send(0, (tag, obj))
        return self._pull(0, lambda t: t == tag)[1]

    def allreduce(self, obj):
        objs = self._xchg_obj(obj, 'allreduce')
        total = objs[0]
        for o in objs[1:]:
            total = total + o
        return total

    def bcast(self, obj, root=0):
        return self._xchg_obj(obj, 'bcast')[root]

    def gather(self, obj, root=0):
        objs = self._xchg_obj(obj, 'gather')
        return objs if self.rank == root else None

    def Barrier(self):
        self._xchg_obj(None, 'barrier')

    # -- vector collectives used by scatter/gather of KronVectorMPI
    def Scatterv(self, sendspec, recvbuf, root=0):
        sendbuf, counts, displs, _ = sendspec
        parts = None
        if self.rank == root:
            flat = np.asarray(sendbuf, dtype=np.float64).reshape(-1)
            parts = [
                flat[int(d):int(d) + int(c)].copy()
                for c, d in zip(counts, displs)
            ]
        parts = self.bcast(parts, root)
        np.asarray(recvbuf)[...] = parts[self.rank].reshape(
            np.asarray(recvbuf).shape)

    def Gatherv(self, sendbuf, recvspec, root=0):
        recvbuf, counts, displs, _ = recvspec
        parts = self._xchg_obj(
            np.array(sendbuf, dtype=np.float64).reshape(-1), 'gatherv')
        if self.rank == root:
            flat = np.asarray(recvbuf).reshape(-1)
            for p, part in enumerate(parts):
                d = int(displs[p])
                flat[d:d + part.size] = part


COMM_WORLD = _Comm()


def _install_world(rank, size, links):
    """Called in a child process by `launch` before the reference is used."""
    global COMM_WORLD
    COMM_WORLD.__init__(rank, size, links)


def _child(rank, size, links, fn, args, result_conn):
    _install_world(rank, size, links)
    try:
        result = fn(*args)
        for q in COMM_WORLD._outbox.values():
            q.join()  # flush the sender threads before this rank exits
        result_conn.send(('ok', result))
    except BaseException as exc:  # surfaced in the parent
        import traceback
        result_conn.send(('err', traceback.format_exc()))
        raise exc


def launch(size, fn, *args):
    """Run fn(*args) on `size` emulated ranks (processes); returns the list of
    per-rank return values.  fn must be a module-level (picklable) function."""
    import multiprocessing as mp
    ctx = mp.get_context('spawn')  # fork after BLAS threads deadlocks
    pipes = {}
    for a in range(size):
        for b in range(a + 1, size):
            pipes[(a, b)] = ctx.Pipe(duplex=True)
    procs, results = [], []
    for r in range(size):
        links = {}
        for p in range(size):
            if p == r:
                continue
            a, b = min(r, p), max(r, p)
            links[p] = pipes[(a, b)][0 if r == a else 1]
        parent_conn, child_conn = ctx.Pipe(duplex=False)
        proc = ctx.Process(target=_child,
                           args=(r, size, links, fn, args, child_conn))
        proc.start()
        procs.append(proc)
        results.append(parent_conn)
    out = []
    for r, conn in enumerate(results):
        status, payload = conn.recv()
        if status != 'ok':
            for p in procs:
                p.terminate()
            raise RuntimeError('rank %d failed:\n%s' % (r, payload))
        out.append(payload)
    for p in procs:
        p.join()
    return out
