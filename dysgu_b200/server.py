"""Parent-side alignment server for fork()ed workers.

`dysgu call -p N` first aligns in the main process (re_map.remap_soft_clips, cluster.pyx:475-478) and only then forks
a multiprocessing.Pool whose workers align again (merge_svs.pyx:396, view.py worker processes).  A process forked after
its parent initialised CUDA cannot use the GPU - a CUDA restriction, reported by the library as DB200_ERR_FORKED
(include/dysgu_b200_align.h, "fork()").  So the parent, which owns the engine, serves such children:

  * the first batch call of a process that may use the GPU starts a daemon thread listening on a Unix socket
    (`ensure_started`); the socket path is a module global, so a later fork()ed child inherits it;
  * a child that finds itself forked after CUDA initialisation (`should_route`) sends its batches - the same numpy CSR
    arrays the C ABI takes - to that socket and gets the result arrays back; the parent's thread runs them through
    dysgu_b200.batch on the parent's engine (ctypes releases the GIL, the pool's parent is blocked in starmap anyway).

Results are identical to a direct call: it is the same engine.  The better fix for dysgu's own pool is
dysgu_b200.patch, which keeps the pool's jobs in the parent and aligns them as a few large device batches; this server is
what makes the unmodified drop-in calls work in ANY forked worker.  DB200_NO_SERVER=1 disables it (children then raise the
library's DB200_ERR_FORKED message).
"""
from __future__ import annotations

import os
import pickle
import socket
import struct
import tempfile
import threading

_path = None          # socket of the process that owns the engine (inherited through fork)
_owner_pid = None
_lock = threading.Lock()
_client = {}          # pid -> connected socket (a forked grandchild reconnects)


def _send(sock, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv(sock):
    hdr = b""
    while len(hdr) < 8:
        chunk = sock.recv(8 - len(hdr))
        if not chunk:
            return None
        hdr += chunk
    n = struct.unpack("<Q", hdr)[0]
    buf = bytearray(n)
    view = memoryview(buf)
    got = 0
    while got < n:
        r = sock.recv_into(view[got:], n - got)
        if r == 0:
            return None
        got += r
    return pickle.loads(bytes(buf))


def _serve_connection(conn):
    from . import batch
    try:
        while True:
            req = _recv(conn)
            if req is None:
                return
            kind, args, kwargs = req
            try:
                if kind == "sw":
                    r = batch.sw_align_batch(*args, **kwargs)
                    out = ("ok", (r.fixed, r.cigar, r.cigar_off, r.status))
                elif kind == "ed":
                    r = batch.ed_align_batch(*args, **kwargs)
                    out = ("ok", (r.fixed, r.loc, r.loc_off, r.path, r.path_off))
                else:
                    out = ("err", "unknown request %r" % (kind,))
            except Exception as exc:   # noqa: BLE001 - travels back to the worker
                out = ("err", "%s: %s" % (type(exc).__name__, exc))
            _send(conn, out)
    finally:
        conn.close()


def _accept_loop(listener):
    while True:
        try:
            conn, _ = listener.accept()
        except OSError:
            return
        threading.Thread(target=_serve_connection, args=(conn,), daemon=True, name="db200-serve").start()


def ensure_started():
    """Called before the first GPU use of a process that can use the GPU: later fork()ed children will find the socket."""
    global _path, _owner_pid
    if os.environ.get("DB200_NO_SERVER") or _owner_pid == os.getpid():
        return
    with _lock:
        if _owner_pid == os.getpid():
            return
        path = os.path.join(tempfile.gettempdir(), "db200-%d-%s.sock" % (os.getpid(), os.urandom(4).hex()))
        listener = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        listener.bind(path)
        listener.listen(64)
        threading.Thread(target=_accept_loop, args=(listener,), daemon=True, name="db200-accept").start()
        import atexit
        atexit.register(lambda p=path, pid=os.getpid(): os.getpid() == pid and os.path.exists(p) and os.unlink(p))
        _path, _owner_pid = path, os.getpid()


def should_route() -> bool:
    """True in a process that was fork()ed after an ancestor initialised CUDA through this library and left a server behind."""
    if _owner_pid is None or _owner_pid == os.getpid() or os.environ.get("DB200_NO_SERVER"):
        return False
    from . import _lib
    return bool(_lib.lib().db200_forked_after_init())


def remote(kind, args, kwargs):
    """Run one batch on the engine of the process this one was forked from."""
    from . import _lib
    pid = os.getpid()
    sock = _client.get(pid)
    if sock is None:
        sock = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        try:
            sock.connect(_path)
        except OSError as exc:
            raise _lib.AlignEngineError("forked worker could not reach the parent's alignment server (%s): %s" % (_path, exc))
        _client.clear()
        _client[pid] = sock
    _send(sock, (kind, args, kwargs))
    rep = _recv(sock)
    if rep is None:
        raise _lib.AlignEngineError("the parent's alignment server closed the connection")
    status, payload = rep
    if status != "ok":
        raise _lib.AlignEngineError("parent-side alignment failed: %s" % payload)
    return payload
