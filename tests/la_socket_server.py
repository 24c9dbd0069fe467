"""A minimal server for the laser-annealing socket protocol, as the FEM side of the reference speaks it
(pymulskips/sockets.py: 12-byte upper-case headers padded with blanks; STATUS -> NEEDINIT | READY | HAVEDATA; INIT;
SENDGEO / SENDFIELD + an int32 array; GETPHASES / GETXGE -> DATAREADY + an int32 array; arrays are
INTEGER(LenZ, LenX*LenY) flattened in Fortran order).  Test infrastructure only."""
import socket
import time

import numpy as np

HDRLEN = 12


def msg(s):
    return s.upper().ljust(HDRLEN).encode()


class Server:
    def __init__(self, shape, timeout=120.0):
        self.shape = shape                       # (LenZ, LenX*LenY)
        self.srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.srv.bind(("127.0.0.1", 0))
        self.port = self.srv.getsockname()[1]
        self.srv.listen(1)
        self.srv.settimeout(timeout)
        self.timeout = timeout
        self.c = None

    def accept(self):
        self.c, _ = self.srv.accept()
        self.c.settimeout(self.timeout)

    def _recv(self, n):
        buf = bytearray()
        while len(buf) < n:
            part = self.c.recv(n - len(buf))
            if not part:
                raise ConnectionError("driver disconnected")
            buf += part
        return bytes(buf)

    def status(self):
        self.c.sendall(msg("status"))
        return self._recv(HDRLEN).decode().strip()

    def wait_status(self, want, tries=100000):
        for _ in range(tries):
            s = self.status()
            if s == want:
                return s
            time.sleep(0.01)
        raise TimeoutError(f"driver never reached {want}")

    def init(self):
        self.c.sendall(msg("init"))

    def send(self, kind, arr):                   # kind: 'geo' | 'field'
        a = np.asarray(arr, dtype=np.int32).reshape(self.shape)
        self.c.sendall(msg("send" + kind))
        self.c.sendall(a.flatten(order="F").tobytes())

    def get(self, kind):                         # kind: 'phases' | 'xge' | 'geo' | 'field'
        self.c.sendall(msg("get" + kind))
        reply = self._recv(HDRLEN).decode().strip()
        assert reply == "DATAREADY", reply
        n = self.shape[0] * self.shape[1]
        return np.frombuffer(self._recv(4 * n), dtype=np.int32).reshape(self.shape, order="F").copy()

    def exit(self):
        try:
            self.c.sendall(msg("exit"))
        except OSError:
            pass

    def close(self):
        for s in (self.c, self.srv):
            try:
                if s:
                    s.close()
            except OSError:
                pass
