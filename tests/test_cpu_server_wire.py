"""-m "not gpu": the nginx-restated CPU server (oracle/nz_cpu_server.c - the wire-level oracle nzproxy is replayed
against in tests/test_gpu_http.py, and bench.py's HTTP reference arm) asked over real sockets beside aiohttp's file
server: the same requests as tests/test_third_party_witness.py, now through request parsing, header framing,
keep-alive and the send loop.  Also what only shows on the wire: pipelined keep-alive requests, HEAD framing, the
`Server: nginx` / `Accept-Ranges` headers (nginx.conf:20), 405 for other methods (nginx.conf:47-49)."""
import asyncio
import http.client
import os
import signal
import socket
import subprocess
import time

import pytest

from oracle import oracle as o

from test_third_party_witness import AGREED, AGREED_AIO, DATA, L, LM, MTIME, Aiohttp, same_answer, with_etag

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST, PATH = "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz"


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.fixture(scope="module")
def wire(tmp_path_factory):
    d = tmp_path_factory.mktemp("cpuwire")
    f = d / "obj.bin"
    f.write_bytes(DATA)
    os.utime(f, (MTIME, MTIME))
    aio = Aiohttp(str(d))
    hx = o.md5(o.cache_key(HOST, PATH)).hex()
    (d / hx).write_bytes(DATA)
    (d / "manifest").write_text(f"{hx} {L} -1 application/octet-stream {aio.etag} - {LM}\n")
    port = free_port()
    p = subprocess.Popen([os.path.join(ROOT, "oracle", "_build", "nz_cpu_server"), "--root", str(d), "--listen", f"127.0.0.1:{port}"], start_new_session=True)
    t0 = time.time()
    while time.time() - t0 < 20:
        try:
            socket.create_connection(("127.0.0.1", port), 0.2).close(); break
        except OSError:
            time.sleep(0.05)
    try:
        yield port, aio
    finally:
        try:
            os.killpg(p.pid, signal.SIGTERM)
        except ProcessLookupError:
            pass


def ask_cpu(port, method, hd, path=PATH):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=20)
    c.putrequest(method, path, skip_host=True, skip_accept_encoding=True)
    c.putheader("Host", HOST)
    for k, v in hd.items():
        c.putheader(k, v)
    c.endheaders()
    r = c.getresponse()
    body = r.read()
    out = (r.status, r.getheader("Content-Range"), r.getheader("Content-Length"), r.getheader("Content-Type", ""), body, dict((k.lower(), v) for k, v in r.getheaders()))
    c.close()
    return out


@pytest.mark.parametrize("method,hd", AGREED + AGREED_AIO, ids=lambda v: str(v))
def test_cpu_server_answers_like_aiohttp_on_the_wire(wire, method, hd):
    port, aio = wire
    hd = with_etag(hd, aio.etag)
    mine = ask_cpu(port, method, hd)
    theirs = aio.ask(method, hd)
    if mine[0] == 416:                                   # nginx sends its stock 416 page (SPEC E1); the comparison is about status + Content-Range
        assert theirs[0] == 416 and (mine[1] or "").endswith(f"*/{L}") and (theirs[1] or "").endswith(f"*/{L}")
        return
    same_answer(mine[:5], theirs, "aiohttp", (method, hd))
    if mine[0] in (200, 206):
        assert mine[5].get("server") == "nginx"
        # nginx's range filter adds Accept-Ranges only where it did NOT serve a range (its `next_filter:` label): 200 yes, 206 no
        assert mine[5].get("accept-ranges") == ("bytes" if mine[0] == 200 else None)
        assert mine[5].get("etag") == aio.etag and mine[5].get("last-modified") == LM


def test_head_has_the_length_and_no_body(wire):
    port, _ = wire
    st, _cr, cl, _ct, body, _h = ask_cpu(port, "HEAD", {})
    assert (st, cl, body) == (200, str(L), b"")


def test_other_methods_are_405_with_the_stock_page(wire):
    port, _ = wire
    st, _cr, _cl, ct, body, hdrs = ask_cpu(port, "POST", {"Content-Length": "0"})
    assert st == 405 and hdrs.get("server") == "nginx" and ct == "text/html" and b"405 Not Allowed" in body


def test_pipelined_keep_alive_requests_come_back_in_order(wire):
    port, _ = wire
    s = socket.create_connection(("127.0.0.1", port), 5)
    reqs = [("bytes=0-9", DATA[0:10]), ("bytes=100-199", DATA[100:200]), ("bytes=-5", DATA[-5:])]
    s.sendall(b"".join(f"GET {PATH} HTTP/1.1\r\nHost: {HOST}\r\nRange: {r}\r\n\r\n".encode() for r, _ in reqs))
    buf = b""
    s.settimeout(5)
    got = []
    while len(got) < len(reqs):
        while b"\r\n\r\n" not in buf:
            buf += s.recv(65536)
        head, _, rest = buf.partition(b"\r\n\r\n")
        assert head.startswith(b"HTTP/1.1 206")
        n = int([l.split(b":")[1] for l in head.split(b"\r\n") if l.lower().startswith(b"content-length:")][0])
        while len(rest) < n:
            rest += s.recv(65536)
        got.append(rest[:n])
        buf = rest[n:]
    s.close()
    assert got == [b for _, b in reqs]


# ---- SPEC P on the wire: raw request heads, answered by status (and the first body bytes) ---------------------------
def raw_exchange(port, head: bytes, timeout=5.0):
    """send one raw head, read one response; returns (status, headers dict, body)"""
    s = socket.create_connection(("127.0.0.1", port), timeout)
    s.settimeout(timeout)
    s.sendall(head)
    buf = b""
    while b"\r\n\r\n" not in buf:
        chunk = s.recv(65536)
        if not chunk:
            break
        buf += chunk
    hd, _, rest = buf.partition(b"\r\n\r\n")
    lines = hd.split(b"\r\n")
    status = int(lines[0].split()[1])
    hdrs = {}
    for ln in lines[1:]:
        k, _, v = ln.partition(b":")
        hdrs[k.decode().lower()] = v.strip().decode("latin-1")
    n = int(hdrs.get("content-length", "0"))
    is_head = head.lstrip(b"\r\n").startswith(b"HEAD ")
    while not is_head and status not in (304,) and len(rest) < n:
        chunk = s.recv(65536)
        if not chunk:
            break
        rest += chunk
    s.close()
    return status, hdrs, rest


def P_CASES(host: str, path: str, etag: str):
    """(raw head, expected status, expected Content-Range or None, expected Connection) - SPEC P rule in the comment"""
    H = f"Host: {host}\r\n".encode()
    G = f"GET {path} HTTP/1.1\r\n".encode()
    E = etag.encode()
    return [
        (G + H + b"\r\n", 200, None, "keep-alive"),
        (f"get {path} HTTP/1.1\r\n".encode() + H + b"\r\n", 400, None, "close"),                               # P1: not a 405
        (f"POST {path} HTTP/1.1\r\n".encode() + H + b"Content-Length: 0\r\n\r\n", 405, None, "keep-alive"),    # V2 sees well-formed methods only
        (f"GET {path} HTTP/2.0\r\n".encode() + H + b"\r\n", 505, None, "close"),                               # P4
        (f"GET {path} HTTP/1.7\r\n".encode() + H + b"\r\n", 200, None, "keep-alive"),                          # P4: behaves as 1.1
        (G + b"\r\n", 400, None, "close"),                                                                     # P8: 1.1 needs Host
        (f"GET {path} HTTP/1.0\r\n".encode() + H + b"Connection: Keep-Alive\r\n\r\n", 200, None, "keep-alive"),   # P9
        (f"GET {path} HTTP/1.0\r\n".encode() + H + b"\r\n", 200, None, "close"),
        (G + H + b"Range: bytes=0-9\r\nRange: bytes=100-199\r\n\r\n", 206, "bytes 0-9/", "keep-alive"),        # P6: first Range line
        (G + H + b"Host: other.example\r\n\r\n", 200, None, "keep-alive"),                                     # P8: first Host line
        (G + H + b"Host: a..b\r\n\r\n", 400, None, "close"),                                                   # P8: every Host line validated
        (G + H + b"If-None-Match: " + E + b"\r\nIf-None-Match: " + E + b"\r\n\r\n", 400, None, "close"),     # P7
        (G + H + b"If-Match: " + E + b"\r\nif-match: \"x\"\r\n\r\n", 400, None, "close"),
        (G + H + b"If-Range: " + E + b"\r\nIf-Range: " + E + b"\r\nRange: bytes=0-0\r\n\r\n", 400, None, "close"),
        (G + H + b"Range : bytes=0-9\r\nX_Y: 1\r\n: v\r\n\r\n", 200, None, "keep-alive"),                      # P5: ignored lines
        (G + H + b"Connection: close\r\nConnection: keep-alive\r\n\r\n", 200, None, "keep-alive"),             # P9: in order
        (G + H + b"Content-Length: 1x\r\n\r\n", 400, None, "close"),                                           # P10
        (b"\r\n" + G + H + b"Range:   bytes=5-6  \r\n\r\n", 206, "bytes 5-6/", "keep-alive"),                  # leading CRLF; spaces trimmed
    ]


def test_request_head_rules_on_the_wire(wire):
    port, aio = wire
    for head, status, crange, conn in P_CASES(HOST, PATH, aio.etag):
        st, hdrs, body = raw_exchange(port, head)
        assert st == status, (head, st, status)
        assert hdrs.get("connection") == conn, (head, hdrs.get("connection"), conn)
        if crange:
            assert hdrs.get("content-range", "").startswith(crange), (head, hdrs.get("content-range"))
        if st == 200 and not head.lstrip(b"\r\n").startswith(b"HEAD"):
            assert body == DATA, head
        if st in (400, 405, 505):
            assert hdrs.get("content-type") == "text/html" and f"<title>{st} ".encode() in body, (head, body[:80])


def test_body_loop_on_an_object_larger_than_the_output_buffers(tmp_path):
    """the pread-32-KiB + send loop (SPEC B1) on 1 MiB + 17 bytes: random single ranges and a three-part request, bodies
    compared with slices of the source (no planner in between)"""
    import random

    from test_third_party_witness import parts_of

    big = bytes(random.Random(5).getrandbits(8) for _ in range((1 << 20) + 17))
    hx = o.md5(o.cache_key(HOST, "/big.bin")).hex()
    (tmp_path / hx).write_bytes(big)
    (tmp_path / "manifest").write_text(f'{hx} {len(big)} -1 application/octet-stream "big" - {LM}\n')
    port = free_port()
    p = subprocess.Popen([os.path.join(ROOT, "oracle", "_build", "nz_cpu_server"), "--root", str(tmp_path), "--listen", f"127.0.0.1:{port}"], start_new_session=True)
    try:
        t0 = time.time()
        while time.time() - t0 < 20:
            try:
                socket.create_connection(("127.0.0.1", port), 0.2).close(); break
            except OSError:
                time.sleep(0.05)
        rnd = random.Random(6)
        n = len(big)
        assert ask_cpu(port, "GET", {}, "/big.bin")[4] == big
        for _ in range(40):
            a = rnd.randrange(n)
            b = min(n - 1, a + rnd.choice([0, 1, 32767, 32768, 32769, 65536, 300_000, n]))
            st, cr, cl, _ct, body, _h = ask_cpu(port, "GET", {"Range": f"bytes={a}-{b}"}, "/big.bin")
            assert (st, cr, cl) == (206, f"bytes {a}-{b}/{n}", str(b - a + 1)) and body == big[a:b + 1], (a, b)
        st, _cr, cl, ct, body, _h = ask_cpu(port, "GET", {"Range": "bytes=0-40000,500000-500000,-70000"}, "/big.bin")
        assert st == 206 and int(cl) == len(body)
        got = parts_of(ct, body)
        assert [c for c, _ in got] == [f"bytes 0-40000/{n}", f"bytes 500000-500000/{n}", f"bytes {n - 70000}-{n - 1}/{n}"]
        assert [b_ for _, b_ in got] == [big[0:40001], big[500000:500001], big[n - 70000:]]
    finally:
        try:
            os.killpg(p.pid, signal.SIGTERM)
        except ProcessLookupError:
            pass


# ---- the whole request path of the CPU server against the pure-Python restatement, head to status ---------------------
def _model_status(raw_head: bytes, known_key: bytes, etag: str):
    """what oracle/spec_py.py says the server must answer for this head: P (head) -> V2 (method) -> K (key) -> hit/miss -> C/R"""
    from oracle import spec_py as sp

    h_ = sp.parse_head(raw_head)
    if h_.error:
        return h_.error
    if not sp.method_gate(h_.method) or h_.method not in ("GET", "HEAD"):
        return 405
    key = sp.cache_key(h_.fields.get("host"), h_.target)
    if key is None:
        return 400
    if key != known_key:
        return 504                                           # the oracle server has no origin: a miss is a gateway time-out
    f = h_.fields
    r = sp.plan(sp.Request(h_.method, h_.minor, f.get("range"), f.get("if-range"), f.get("if-none-match"), f.get("if-match"), f.get("if-modified-since"),
                           f.get("if-unmodified-since"), f.get("accept-encoding")), sp.Stored(L, "application/octet-stream", etag, LM, None, -1, False, True), 7)
    return r.status


def test_whole_request_path_against_the_python_model(wire):
    import random

    from oracle import spec_py as sp

    port, aio = wire
    known = sp.cache_key(HOST, PATH)
    rnd = random.Random(11)
    lines = [f"GET {PATH} HTTP/1.1", f"HEAD {PATH} HTTP/1.1", f"GET {PATH} HTTP/1.0", f"GET {PATH}", f"get {PATH} HTTP/1.1", f"POST {PATH} HTTP/1.1", f"XGETX {PATH} HTTP/1.1",
             f"GET {PATH} HTTP/2.0", f"GET {PATH} HTTP/1.9", f"GET /other HTTP/1.1", f"GET /a/../..{PATH} HTTP/1.1", f"GET //~kriz/./cifar-100-python.tar.gz?x=1 HTTP/1.1",
             f"GET http://{HOST}{PATH} HTTP/1.1", f"GET {PATH} HTTP/1.1 x", f"GET  {PATH}   HTTP/1.1"]
    names = ["Host", "host", "Range", "If-Range", "If-Match", "If-None-Match", "If-Modified-Since", "If-Unmodified-Since", "Connection", "Content-Length", "X_Bad", "Accept-Encoding"]
    values = [HOST, HOST.upper() + ":80", "other.example", "a..b", "bytes=0-9", "bytes=5-", "bytes=-0", "bytes=20000-", "bytes=0-1,5-6", aio.etag, '"x"', "*", LM, "junk", "close", "0", "7", "1x", "gzip"]
    for _ in range(400):
        rl = rnd.choice(lines)
        hdrs = [(rnd.choice(names), rnd.choice(values)) for _ in range(rnd.randrange(0, 5))]
        if rnd.random() < 0.8:
            hdrs.insert(0, ("Host", HOST))
        raw = (rl + "\r\n" + "".join(f"{n}: {v}\r\n" for n, v in hdrs)).encode()
        want = _model_status(raw, known, aio.etag)
        got = raw_exchange(port, raw + b"\r\n")[0]
        assert got == want, (raw, got, want)
