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
{"cmd": "shutdown" | "ping"}; reply {"rc": exit code, "out": / "err": what ./main would have printed on stdout / stderr}.  Requests are served
one at a time (a solve takes ~5 ms; the 32 PATH_MAP workers simply queue on the socket).
"""
import argparse
import io
import json
import os
import socket
import sys
from contextlib import redirect_stderr, redirect_stdout


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
            conn, _ = srv.accept()
            with conn:
                conn.settimeout(10.0)                        # a silent or vanished client must not stall the queue
                try:
                    if _serve_one(conn, towr_cli, device):
                        return
                except OSError:
                    continue
    finally:
        srv.close()
        try:
            os.remove(sock_path)
        except OSError:
            pass


def _serve_one(conn, towr_cli, device):
    """one request on an accepted connection; True when it was the shutdown command."""
    buf = b""
    while not buf.endswith(b"\n"):
        chunk = conn.recv(65536)
        if not chunk:
            break
        buf += chunk
    try:
        req = json.loads(buf.decode())
    except ValueError:
        conn.sendall(b'{"rc": 3, "out": "bad request"}\n')
        return False
    if req.get("cmd") == "shutdown":
        conn.sendall(b'{"rc": 0, "out": "bye"}\n')
        return True
    if req.get("cmd") == "ping":
        conn.sendall(b'{"rc": 0, "out": "pong"}\n')
        return False
    out, err = io.StringIO(), io.StringIO()
    try:
        with redirect_stdout(out), redirect_stderr(err):
            rc = towr_cli.towr_main(list(req.get("argv", [])), cwd=req.get("cwd", "."), device=device)
    except Exception as e:                       # the daemon outlives a bad request
        rc = 3
        err.write("qtos: %s\n" % e)
    conn.sendall((json.dumps({"rc": int(rc), "out": out.getvalue(), "err": err.getvalue()}) + "\n").encode())
    return False


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--socket", default=None)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    serve(a.socket, a.device, ready=lambda: (sys.stdout.write("qtos_b200 daemon ready\n"), sys.stdout.flush()))
