"""Long-lived solver process behind the `./main` boundary.

A fresh `./main` process spends 0.8 of its 1.1 s creating a CUDA context and compiling the shape;
the solve itself takes milliseconds.  QTOS calls the local planner once per replanning step and up to 32 times in
parallel from PATH_MAP (ref: scripts/main.py:48-50,90-92; QTOS/generateHeightField.py:344-404), always through a
command line, so the zero-edit way to remove that cost is a daemon that keeps the context, the compiled shapes and
the uploaded heightfields, and a thin client behind the same command line:

    python -m qtos_b200.serve [--socket PATH] [--device N]        the daemon (foreground)
    shim/docker exec <id> ./main <flags>                           asks the daemon when its socket answers,
                                                                   runs the native `main` binary otherwise

Protocol: one JSON object per connection, newline-terminated -- request {"argv": [...], "cwd": "..."} or
{"cmd": "shutdown" | "ping"}; reply {"rc": exit code, "out": / "err": what ./main would have printed on stdout / stderr}.  Every connection
that is already waiting when the daemon comes back to its socket is taken in the same round, and the ./main requests of a
round are solved as ONE batch (towr_cli.towr_main_many): the 32 PATH_MAP workers cost one launch sequence, not 32.
"""
import argparse
import json
import os
import socket
import sys


def default_socket():
    root = os.environ.get("QTOS_SHIM_ROOT", os.path.join(os.path.expanduser("~"), ".qtos_b200", "towr"))
    return os.path.join(root, "qtos.sock")


def request(sock_path, obj, timeout=60.0):
    """client side: returns the reply dict, or None when no daemon answers on sock_path."""
    try:
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.settimeout(timeout)
        s.connect(sock_path)
    except OSError:
        return None
    with s:
        s.sendall((json.dumps(obj) + "\n").encode())
        buf = b""
        while not buf.endswith(b"\n"):
            chunk = s.recv(65536)
            if not chunk:
                break
            buf += chunk
    return json.loads(buf.decode()) if buf else None


def serve(sock_path=None, device=0, ready=None):
    from . import towr_cli
    sock_path = sock_path or default_socket()
    os.makedirs(os.path.dirname(sock_path), exist_ok=True)
    if os.path.exists(sock_path):
        if request(sock_path, {"cmd": "ping"}, timeout=2.0) is not None:
            raise RuntimeError("a daemon already answers on %s" % sock_path)
        os.remove(sock_path)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(sock_path)
    srv.listen(64)
    if ready is not None:
        ready()
    try:
        while True:
            conns = [srv.accept()[0]]
            srv.setblocking(False)
            try:                                             # everything that queued up meanwhile joins this round
                while len(conns) < MAX_ROUND:
                    conns.append(srv.accept()[0])
            except (BlockingIOError, InterruptedError):
                pass
            finally:
                srv.setblocking(True)
            if _serve_round(conns, towr_cli, device):
                return
    finally:
        srv.close()
        try:
            os.remove(sock_path)
        except OSError:
            pass


MAX_ROUND = 64


def _read_request(conn):
    buf = b""
    while not buf.endswith(b"\n"):
        chunk = conn.recv(65536)
        if not chunk:
            break
        buf += chunk
    req = json.loads(buf.decode())
    if not isinstance(req, dict):
        raise ValueError("request is not a JSON object")
    return req


def _reply(conn, obj):
    try:
        conn.sendall((json.dumps(obj) + "\n").encode())
    except OSError:
        pass                                                 # the client went away: nobody to tell


def _serve_round(conns, towr_cli, device):
    """the requests of the connections accepted in one round; True when one of them was the shutdown command."""
    shutdown = False
    jobs = []
    for conn in conns:
        conn.settimeout(10.0)                                # a silent or vanished client must not stall the queue
        try:
            req = _read_request(conn)
        except (OSError, ValueError):
            _reply(conn, {"rc": 3, "out": "bad request", "err": ""})
            conn.close()
            continue
        if req.get("cmd") == "shutdown":
            _reply(conn, {"rc": 0, "out": "bye"}); conn.close(); shutdown = True
        elif req.get("cmd") == "ping":
            _reply(conn, {"rc": 0, "out": "pong"}); conn.close()
        else:
            argv = req.get("argv", [])
            jobs.append((conn, [str(a) for a in argv] if isinstance(argv, list) else [], str(req.get("cwd", "."))))
    if jobs:
        try:
            results = towr_cli.towr_main_many([(argv, cwd) for _, argv, cwd in jobs], device=device)
        except Exception as e:                               # the daemon outlives a bad request
            results = [(3, "", "qtos: %s\n" % e)] * len(jobs)
        for (conn, _, _), (rc, out, err) in zip(jobs, results):
            _reply(conn, {"rc": int(rc), "out": out, "err": err})
            conn.close()
    return shutdown


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--socket", default=None)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    serve(a.socket, a.device, ready=lambda: (sys.stdout.write("qtos_b200 daemon ready\n"), sys.stdout.flush()))
