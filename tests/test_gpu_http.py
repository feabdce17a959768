"""-m gpu: HTTP-level differential replay (SURVEY.md §4 item 3, §8c parity definition).

The same request log is sent to
  * nzproxy          — the product: C++ front end over the C ABI, bodies from HBM, and
  * nz_cpu_server    — the oracle's nginx-restated CPU server (pread 32 KiB + send),
and status, the header subset {Content-Length, Content-Range, Content-Type, ETag,
Last-Modified, Accept-Ranges, Content-Encoding} and the body SHA-256 must agree
(multipart boundaries normalised; Date/Server/Connection ignored).  nzproxy starts
cold, so the first GET of every object exercises the miss -> origin fetch -> fill path.
"""
import hashlib
import http.client
import http.server
import os
import re
import signal
import socket
import subprocess
import threading
import time

import pytest

from nezha_b200 import workloads as wl
from oracle import oracle as o

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LM = "Wed, 07 Nov 2018 15:20:01 GMT"
SUBSET = ["content-length", "content-range", "content-type", "etag", "last-modified", "accept-ranges", "content-encoding"]


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def wait_port(port, timeout=60):
    t0 = time.time()
    while time.time() - t0 < timeout:
        try:
            socket.create_connection(("127.0.0.1", port), 0.2).close()
            return
        except OSError:
            time.sleep(0.1)
    raise RuntimeError(f"port {port} never opened")


OBJECTS = {
    # path -> (host, bytes, content-type, etag, encoding)
    "/~kriz/cifar-100-python.tar.gz": ("www.cs.toronto.edu", wl.synth_bytes(3_000_017, 1).tobytes(), "application/x-gzip", '"5be2f2c5-a12c15d"', None),
    "/exdb/mnist/train-images-idx3-ubyte.gz": ("yann.lecun.com", wl.synth_bytes(70_001, 2).tobytes(), "application/octet-stream", None, None),
    "/bucket/big.bin": ("storage.googleapis.com", wl.synth_bytes(20 << 20, 3).tobytes(), "application/octet-stream", '"big"', None),
    "/bucket/empty": ("storage.googleapis.com", b"", "application/octet-stream", '"e"', None),
    "/zips/text.json.gz": ("images.cocodataset.org", wl.gzip_member(wl.text_like(500_000, 9)), "application/json", '"gz1"', "gzip"),
    # an origin that does not say Accept-Ranges: nginx serves the cached copy whole, whatever Range says (SPEC R0)
    "/bucket/noranges.bin": ("storage.googleapis.com", wl.synth_bytes(300_001, 4).tobytes(), "application/octet-stream", '"nr"', None),
}
NO_ACCEPT_RANGES = {"/bucket/noranges.bin"}


class Origin(http.server.BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.0"
    hits = {}

    def do_GET(self):
        path = self.path.split("?")[0]
        Origin.hits[path] = Origin.hits.get(path, 0) + 1
        ent = OBJECTS.get(path)
        if not ent:
            self.send_response(404)
            self.send_header("Content-Length", "0")
            self.end_headers()
            return
        _, data, ctype, etag, enc = ent
        self.send_response(200)
        self.send_header("Content-Type", ctype)
        self.send_header("Content-Length", str(len(data)))
        if etag:
            self.send_header("ETag", etag)
        self.send_header("Last-Modified", LM)
        if path not in NO_ACCEPT_RANGES:
            self.send_header("Accept-Ranges", "bytes")        # what GCS / S3 / Apache answer; nginx needs it to serve ranges from the cache
        if enc:
            self.send_header("Content-Encoding", enc)
        self.end_headers()
        self.wfile.write(data)

    def log_message(self, *a):
        pass


@pytest.fixture(scope="module")
def servers(tmp_path_factory):
    d = tmp_path_factory.mktemp("cache")
    lines = []
    for path, (host, data, ctype, etag, enc) in OBJECTS.items():
        key = o.cache_key(host, path)
        hx = o.md5(key).hex()
        (d / hx).write_bytes(data)
        raw = len(o.gunzip(data)) if enc else -1
        lines.append(f"{hx} {len(data)} {raw} {ctype} {etag or '-'} {enc or '-'}{',noranges' if path in NO_ACCEPT_RANGES else ''} {LM}")
    (d / "manifest").write_text("\n".join(lines) + "\n")
    origin_port, gpu_port, cpu_port = free_port(), free_port(), free_port()
    httpd = http.server.ThreadingHTTPServer(("127.0.0.1", origin_port), Origin)
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    ups = []
    for h in {v[0] for v in OBJECTS.values()}:
        ups += ["--upstream", f"{h}=127.0.0.1:{origin_port}"]
    procs = [
        subprocess.Popen([os.path.join(ROOT, "nezha_b200", "_build", "nzproxy"), "--listen", f"127.0.0.1:{gpu_port}", "--arena-gib", "1",
                          "--max-conns", "32", "--window-mib", "1", "--first-window-kib", "64", "--verify"] + ups, start_new_session=True),
        subprocess.Popen([os.path.join(ROOT, "oracle", "_build", "nz_cpu_server"), "--root", str(d), "--listen", f"127.0.0.1:{cpu_port}"],
                         start_new_session=True),
    ]
    try:
        wait_port(cpu_port)
        wait_port(gpu_port, 120)
        yield gpu_port, cpu_port
    finally:
        for p in procs:
            try:
                os.killpg(p.pid, signal.SIGTERM)
            except ProcessLookupError:
                pass
        httpd.shutdown()


def fetch(port, method, host, target, headers):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
    c.putrequest(method, target, skip_host=True, skip_accept_encoding=True)
    if host is not None:
        c.putheader("Host", host)
    for k, v in headers.items():
        c.putheader(k, v)
    c.endheaders()
    r = c.getresponse()
    body = r.read()
    hd = {k.lower(): v for k, v in r.getheaders()}
    c.close()
    return r.status, hd, body


def normalise(hd, body):
    ct = hd.get("content-type", "")
    m = re.match(r"multipart/byteranges; boundary=(\d{20})$", ct)
    if m:
        b = m.group(1).encode()
        body = body.replace(b, b"B" * 20)
        hd = dict(hd)
        hd["content-type"] = "multipart/byteranges; boundary=" + "B" * 20
    return {k: hd.get(k) for k in SUBSET}, hashlib.sha256(body).hexdigest(), len(body)


LOG = []
for _path, (_host, _data, *_r) in OBJECTS.items():
    n = len(_data)
    LOG += [("GET", _host, _path, {}), ("HEAD", _host, _path, {}), ("GET", _host, _path, {"Range": "bytes=0-0"}),
            ("GET", _host, _path, {"Range": f"bytes={max(n - 5, 0)}-"}), ("GET", _host, _path, {"Range": "bytes=-7"}),
            ("GET", _host, _path, {"Range": f"bytes={n}-"}), ("GET", _host, _path + "?alt=media", {"Range": "bytes=1-3,10-12,100-"}),
            ("GET", _host, _path, {"If-Modified-Since": LM}), ("GET", _host, _path, {"If-None-Match": "*"}),
            ("GET", _host, _path, {"If-Match": '"nope"'}), ("GET", _host.upper() + ":80", _path, {"Range": "bytes=5-4"}),
            ("GET", _host, "//" + _path.lstrip("/").replace("/", "/./", 1), {"Range": "bytes=65530-65540"})]
LOG += [
    ("GET", "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz", {"Range": "bytes=7-2000000", "If-Range": '"5be2f2c5-a12c15d"'}),
    ("GET", "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz", {"Range": "bytes=7-2000000", "If-Range": '"other"'}),
    ("GET", "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz", {"Range": "bytes=1000000-1000099,5-70000, 2999999-"}),
    ("GET", "storage.googleapis.com", "/bucket/big.bin", {"Range": "bytes=1048575-9437184"}),
    ("GET", "storage.googleapis.com", "/bucket/big.bin", {"Range": "bytes=0-20971519"}),
    ("POST", "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz", {}),
    ("GET", "www.cs.toronto.edu", "/../etc/passwd", {}),
    ("GET", "images.cocodataset.org", "/zips/text.json.gz", {"Accept-Encoding": "gzip"}),
    # SPEC section Z (what the second restatement changed), over the wire against both servers
    ("GET", "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz", {"If-Unmodified-Since": "not a date", "Range": "bytes=0-9"}),      # Z1: 412
    ("GET", "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz", {"If-Unmodified-Since": LM, "Range": "bytes=0-9"}),                # 206
    ("GET", "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz", {"If-Unmodified-Since": "Wed, 00 Nov 2018 15:20:01 GMT"}),         # Z4: day 00 parses, earlier than LM: 412
    ("GET", "storage.googleapis.com:80/x", "/bucket/big.bin", {}),                                                                    # Z3: '/' after the port: 400
    ("GET", "storage.googleapis.com", "/bucket/noranges.bin", {"Range": "bytes=5-9", "If-Range": '"nr"'}),                             # Z5: still the whole object
]


def test_differential_replay(servers):
    gpu, cpu = servers
    for method, host, target, hdrs in LOG:
        sg, hg, bg = fetch(gpu, method, host, target, hdrs)
        sc, hc, bc = fetch(cpu, method, host, target, hdrs)
        assert sg == sc, (method, host, target, hdrs, sg, sc)
        assert normalise(hg, bg) == normalise(hc, bc), (method, host, target, hdrs)
        assert hg.get("server") == "nginx"
    # every object was fetched from the origin exactly once (miss -> fill -> hits)
    for path in OBJECTS:
        assert Origin.hits.get(path) == 1, (path, Origin.hits)


def test_bodies_match_source_and_keepalive(servers):
    gpu, _ = servers
    c = http.client.HTTPConnection("127.0.0.1", gpu, timeout=60)
    for path, (host, data, *_r) in OBJECTS.items():
        for _ in range(2):                                   # same connection: keep-alive
            c.request("GET", path, headers={"Host": host})
            r = c.getresponse()
            assert r.status == 200 and r.read() == data
    c.request("GET", "/bucket/big.bin", headers={"Host": "storage.googleapis.com", "Range": "bytes=12345-7654321"})
    r = c.getresponse()
    assert r.status == 206 and r.read() == OBJECTS["/bucket/big.bin"][1][12345:7654322]
    c.close()


def test_origin_404_is_not_cached(servers):
    gpu, _ = servers
    s1, _, _ = fetch(gpu, "GET", "yann.lecun.com", "/nope", {})
    s2, _, _ = fetch(gpu, "GET", "yann.lecun.com", "/nope", {})
    assert s1 == 404 and s2 == 404 and Origin.hits.get("/nope") == 2      # proxy_cache_valid 200 only


# ---- fill-side policy: integrity header, durable tier, cache manager ----
def _start_proxy(port, origin_port, extra):
    ups = []
    for h in ("www.cs.toronto.edu", "yann.lecun.com", "storage.googleapis.com", "images.cocodataset.org"):
        ups += ["--upstream", f"{h}=127.0.0.1:{origin_port}"]
    log = open(os.path.join(ROOT, "gpurun_out", f"nzproxy_{port}.log"), "wb") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else subprocess.DEVNULL
    p = subprocess.Popen([os.path.join(ROOT, "nezha_b200", "_build", "nzproxy"), "--listen", f"127.0.0.1:{port}", "--max-conns", "16",
                          "--window-mib", "1"] + ups + extra, start_new_session=True, stderr=log)
    wait_port(port, 120)
    return p


def _stop(p):
    try:
        os.killpg(p.pid, signal.SIGTERM)
    except ProcessLookupError:
        pass
    p.wait(timeout=20)


class PolicyOrigin(http.server.BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.0"
    hits = {}
    objects = {}      # path -> (bytes, extra headers)

    def do_GET(self):
        path = self.path.split("?")[0]
        PolicyOrigin.hits[path] = PolicyOrigin.hits.get(path, 0) + 1
        ent = PolicyOrigin.objects.get(path)
        if not ent:
            self.send_response(404); self.send_header("Content-Length", "0"); self.end_headers(); return
        data, extra = ent
        self.send_response(200)
        self.send_header("Content-Type", "application/octet-stream")
        self.send_header("Content-Length", str(len(data)))
        self.send_header("ETag", '"%s"' % hashlib.md5(data).hexdigest())
        self.send_header("Accept-Ranges", "bytes")
        for k, v in extra.items():
            self.send_header(k, v)
        self.end_headers()
        self.wfile.write(data)

    def log_message(self, *a):
        pass


@pytest.fixture(scope="module")
def policy_origin():
    port = free_port()
    httpd = http.server.ThreadingHTTPServer(("127.0.0.1", port), PolicyOrigin)
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    yield port
    httpd.shutdown()


def test_gcs_crc32c_header_is_verified_on_fill(policy_origin):
    good = wl.synth_bytes(1_234_567, 70).tobytes()
    bad = wl.synth_bytes(200_000, 71).tobytes()
    PolicyOrigin.objects["/b/good"] = (good, {"x-goog-hash": "crc32c=" + o.crc32c_header(o.crc32c(good))})
    PolicyOrigin.objects["/b/bad"] = (bad, {"x-goog-hash": "crc32c=" + o.crc32c_header(o.crc32c(bad) ^ 0x10) + ",md5=AAAA"})
    port = free_port()
    import tempfile
    alog = tempfile.mktemp(prefix="nz_access_", suffix=".log")
    p = _start_proxy(port, policy_origin, ["--arena-gib", "0.5", "--access-log", alog])
    try:
        for _ in range(2):
            s, _, b = fetch(port, "GET", "storage.googleapis.com", "/b/good", {"User-Agent": "Wget/1.19", "Referer": "http://example/"})
            assert s == 200 and b == good
        assert PolicyOrigin.hits["/b/good"] == 1
        s, _, b = fetch(port, "GET", "storage.googleapis.com", "/b/good", {"Range": "bytes=10-19"})
        assert s == 206
        s, _, _ = fetch(port, "HEAD", "storage.googleapis.com", "/b/good", {})
        assert s == 200
        # access_log ... main (nginx.conf:16-19)
        import re
        lines = open(alog).read().splitlines()
        assert len(lines) == 4
        pat = re.compile(r'^127\.0\.0\.1 - - \[\d{2}/[A-Z][a-z]{2}/\d{4}:\d{2}:\d{2}:\d{2} [+-]\d{4}\] "(GET|HEAD) /b/good HTTP/1\.1" (\d{3}) (\d+) "([^"]*)" "([^"]*)" "-"$')
        got = [pat.match(l).groups() for l in lines]
        assert [(g[0], g[1], int(g[2])) for g in got] == [("GET", "200", len(good)), ("GET", "200", len(good)), ("GET", "206", 10), ("HEAD", "200", 0)]
        assert got[0][3] == "http://example/" and got[0][4] == "Wget/1.19" and got[2][3] == "-"
        s1, _, _ = fetch(port, "GET", "storage.googleapis.com", "/b/bad", {})
        s2, _, _ = fetch(port, "GET", "storage.googleapis.com", "/b/bad", {})
        assert s1 == 502 and s2 == 502 and PolicyOrigin.hits["/b/bad"] == 2      # never admitted (NZ_ECORRUPT at commit)
    finally:
        _stop(p)


def test_pv_write_through_and_cold_start(policy_origin, tmp_path):
    pv = tmp_path / "pv"
    pv.mkdir()
    objs = {f"/pv/obj{i}": wl.synth_bytes(2_000_000 + 333 * i, 80 + i).tobytes() for i in range(4)}
    for k, v in objs.items():
        PolicyOrigin.objects[k] = (v, {"Last-Modified": LM})
    port = free_port()
    p = _start_proxy(port, policy_origin, ["--arena-gib", "0.5", "--pv", str(pv)])
    try:
        for k, v in objs.items():
            s, _, b = fetch(port, "GET", "yann.lecun.com", k, {})
            assert s == 200 and b == v
    finally:
        _stop(p)
    files = sorted(os.listdir(pv))
    assert "index.tsv" in files and len([f for f in files if len(f) == 32]) == 4 and not [f for f in files if f.endswith(".part")]
    before = dict(PolicyOrigin.hits)
    port2 = free_port()
    p2 = _start_proxy(port2, 1, ["--arena-gib", "0.5", "--pv", str(pv)])          # origin unreachable on purpose (port 1)
    try:
        for k, v in objs.items():
            s, hd, b = fetch(port2, "GET", "yann.lecun.com", k, {"Range": "bytes=1000-1999999"})
            assert s == 206 and b == v[1000:2000000] and hd.get("last-modified") == LM
        assert PolicyOrigin.hits == before                                         # served from the reloaded PV copy
        s, _, _ = fetch(port2, "GET", "yann.lecun.com", "/pv/never-seen", {})
        assert s == 502
    finally:
        _stop(p2)
    # a corrupted PV file is rejected at reload (CRC-32C recorded at fill)
    victim = [f for f in files if len(f) == 32][0]
    raw = bytearray((pv / victim).read_bytes()); raw[1000] ^= 1; (pv / victim).write_bytes(bytes(raw))
    port3 = free_port()
    p3 = _start_proxy(port3, 1, ["--arena-gib", "0.5", "--pv", str(pv)])
    try:
        codes = sorted(fetch(port3, "GET", "yann.lecun.com", k, {})[0] for k in objs)
        assert codes == [200, 200, 200, 502]
    finally:
        _stop(p3)


def test_inactive_purge_and_lru_under_pressure(policy_origin):
    """SPEC H1 (absent key -> MISS -> fill), H2 (`proxy_cache_valid 200 1d`: --valid-s, an expired entry is a MISS again),
    H3 (`inactive=60m`: --inactive-s purge), plus LRU eviction when the arena is full (nginx has no max_size; HBM does)."""
    objs = {f"/lru/obj{i}": wl.synth_bytes(24 << 20, 90 + i).tobytes() for i in range(5)}     # 5 x 24 MiB into a 64 MiB arena
    for k, v in objs.items():
        PolicyOrigin.objects[k] = (v, {})
    port = free_port()
    p = _start_proxy(port, policy_origin, ["--arena-gib", "0.0625", "--inactive-s", "3"])
    try:
        for k, v in objs.items():
            s, _, b = fetch(port, "GET", "www.cs.toronto.edu", k, {})
            assert s == 200 and hashlib.sha256(b).digest() == hashlib.sha256(v).digest()    # the arena holds 2: older ones were evicted to make room
        last = "/lru/obj4"
        s, _, b = fetch(port, "GET", "www.cs.toronto.edu", last, {"Range": "bytes=0-9"})
        assert s == 206 and b == objs[last][:10] and PolicyOrigin.hits[last] == 1          # most recent one is still cached
        s, _, _ = fetch(port, "GET", "www.cs.toronto.edu", "/lru/obj0", {"Range": "bytes=0-9"})
        assert PolicyOrigin.hits["/lru/obj0"] == 2                                          # the oldest was evicted and refetched
        time.sleep(6)                                                                       # inactive=3 s: everything idles out
        s, _, _ = fetch(port, "GET", "www.cs.toronto.edu", last, {"Range": "bytes=0-9"})
        assert s == 206 and PolicyOrigin.hits[last] == 2
    finally:
        _stop(p)


def test_front_end_shards_across_two_gpus(policy_origin):
    """nzproxy --devices 0,1: one zone per GPU, every GET routed to the shard its key hashes to."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from nezha_b200 import shard
    objs = {f"/two/obj{i}": wl.synth_bytes(3_000_000 + 17 * i, 120 + i).tobytes() for i in range(12)}
    owners = set()
    for k, v in objs.items():
        PolicyOrigin.objects[k] = (v, {})
        owners.add(shard.shard_of(o.md5(o.cache_key("images.cocodataset.org", k)), 2))
    assert owners == {0, 1}                                  # the log really exercises both GPUs
    port = free_port()
    p = _start_proxy(port, policy_origin, ["--arena-gib", "0.5", "--devices", "0,1"])
    try:
        for rnd in range(2):
            for k, v in objs.items():
                s, _, b = fetch(port, "GET", "images.cocodataset.org", k, {"Range": "bytes=100-2999999"} if rnd else {})
                assert s == (206 if rnd else 200) and b == (v[100:3000000] if rnd else v), k
        assert all(PolicyOrigin.hits[k] == 1 for k in objs)
    finally:
        _stop(p)


# ---- https upstream of the GCS vhost (nginx.conf:113) and the origins' MD5 validators ----
def test_tls_upstream_and_md5_validators(policy_origin, tmp_path):
    """`proxy_pass https://storage.googleapis.com`: the fill goes over TLS with nginx's defaults (no verification,
    no SNI); Content-MD5 (Azure Blob, RFC 1864) and x-goog-hash md5= are checked against the bytes that arrived."""
    import base64
    import ssl
    crt, key = tmp_path / "o.crt", tmp_path / "o.key"
    r = subprocess.run(["openssl", "req", "-x509", "-newkey", "rsa:2048", "-nodes", "-days", "2", "-subj", "/CN=localhost",
                        "-keyout", str(key), "-out", str(crt)], capture_output=True)
    if r.returncode != 0:
        pytest.skip("openssl CLI not usable here")
    tport = free_port()
    httpd = http.server.ThreadingHTTPServer(("127.0.0.1", tport), PolicyOrigin)
    cx = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
    cx.load_cert_chain(str(crt), str(key))
    httpd.socket = cx.wrap_socket(httpd.socket, server_side=True)
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    obj = wl.synth_bytes(3_333_333, 90).tobytes()
    md5_b64 = base64.b64encode(hashlib.md5(obj).digest()).decode()
    wrong = base64.b64encode(hashlib.md5(obj + b"x").digest()).decode()
    PolicyOrigin.objects["/tls/obj"] = (obj, {"x-goog-hash": "crc32c=" + o.crc32c_header(o.crc32c(obj)) + ",md5=" + md5_b64})
    PolicyOrigin.objects["/tls/azure-ok"] = (obj, {"Content-MD5": md5_b64})
    PolicyOrigin.objects["/tls/azure-bad"] = (obj, {"Content-MD5": wrong})
    PolicyOrigin.objects["/tls/goog-md5-bad"] = (obj, {"x-goog-hash": "md5=" + wrong})
    port = free_port()
    log = open(os.path.join(ROOT, "gpurun_out", f"nzproxy_{port}.log"), "wb") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else subprocess.DEVNULL
    p = subprocess.Popen([os.path.join(ROOT, "nezha_b200", "_build", "nzproxy"), "--listen", f"127.0.0.1:{port}", "--max-conns", "16", "--window-mib", "1",
                          "--arena-gib", "0.5", "--upstream", f"storage.googleapis.com=https://127.0.0.1:{tport}",
                          "--upstream", f"www.cs.toronto.edu=127.0.0.1:{policy_origin}"], start_new_session=True, stderr=log)
    wait_port(port, 120)
    try:
        before = PolicyOrigin.hits.get("/tls/obj", 0)
        for _ in range(2):
            s, hd, b = fetch(port, "GET", "storage.googleapis.com", "/tls/obj", {})
            assert s == 200 and b == obj
        assert PolicyOrigin.hits["/tls/obj"] == before + 1                       # second GET is a hit from HBM
        s, _, b = fetch(port, "GET", "storage.googleapis.com", "/tls/obj", {"Range": "bytes=1000000-1000099"})
        assert s == 206 and b == obj[1000000:1000100]
        s, _, b = fetch(port, "GET", "storage.googleapis.com", "/tls/azure-ok", {})
        assert s == 200 and b == obj
        for path in ("/tls/azure-bad", "/tls/goog-md5-bad"):
            s1, _, _ = fetch(port, "GET", "storage.googleapis.com", path, {})
            s2, _, _ = fetch(port, "GET", "storage.googleapis.com", path, {})
            assert s1 == 502 and s2 == 502 and PolicyOrigin.hits[path] == 2      # never admitted
        # a plain-http vhost next to it still works, and a TLS port spoken to in plain text fails closed (502)
        PolicyOrigin.objects["/plain"] = (b"hello" * 1000, {})
        s, _, b = fetch(port, "GET", "www.cs.toronto.edu", "/plain", {})
        assert s == 200 and b == b"hello" * 1000
    finally:
        _stop(p)
        httpd.shutdown()


# ---- validity from upstream headers (SPEC H2) at the wire -------------------------------------
def test_upstream_validity_headers_H2(policy_origin):
    body = {k: wl.synth_bytes(150_000 + 7 * i, 300 + i).tobytes() for i, k in enumerate(
        ["maxage1", "nostore", "xae0", "xae_over_cc", "cc_then_xae", "cookie", "expired", "smaxage", "plain"])}
    hdr = {
        "maxage1": {"Cache-Control": "public, max-age=1"},
        "nostore": {"Cache-Control": "no-store"},
        "xae0": {"X-Accel-Expires": "0"},
        "xae_over_cc": {"X-Accel-Expires": "3600", "Cache-Control": "no-cache"},      # arrival order: XAE first -> CC ignored (H2d)
        "cc_then_xae": {"Cache-Control": "private", "X-Accel-Expires": "3600"},       # CC first -> sticky uncacheable
        "cookie": {"Set-Cookie": "sid=1"},
        "expired": {"Expires": "Wed, 07 Nov 2018 00:00:00 GMT"},
        "smaxage": {"Cache-Control": "max-age=0, s-maxage=3600"},                     # s-maxage is read first
        "plain": {},
    }
    for k in body:
        PolicyOrigin.objects["/v/" + k] = (body[k], hdr[k])
    port = free_port()
    p = _start_proxy(port, policy_origin, ["--arena-gib", "0.5"])
    try:
        def get(k):
            s, hd, b = fetch(port, "GET", "storage.googleapis.com", "/v/" + k, {})
            assert s == 200 and b == body[k], (k, s, len(b))
            return hd
        for k in body:
            PolicyOrigin.hits.pop("/v/" + k, None)
            get(k); get(k)
        cached = {k: PolicyOrigin.hits["/v/" + k] == 1 for k in body}
        assert cached == {"maxage1": True, "nostore": False, "xae0": False, "xae_over_cc": True, "cc_then_xae": False, "cookie": False,
                          "expired": False, "smaxage": True, "plain": True}, cached
        # the oracle's restatement of the same rule agrees, header for header
        now = int(time.time())
        for k in body:
            assert o.admission(200, list(hdr[k].items()), now)[0] == cached[k], k
        # an uncacheable response is relayed with the origin's own headers
        assert get("cookie").get("set-cookie") == "sid=1" and get("nostore").get("cache-control") == "no-store"
        time.sleep(2.3)                                        # max-age=1 has run out: refetched once, then cached again
        get("maxage1"); get("maxage1")
        assert PolicyOrigin.hits["/v/maxage1"] == 2
        get("plain")
        assert PolicyOrigin.hits["/v/plain"] == 1              # proxy_cache_valid 200 1d still covers it
    finally:
        _stop(p)


# ---- SPEC G1-G4 at the wire: nzproxy --gunzip vs the restated CPU server --gunzip -------------------------
def test_gunzip_switch_differential(policy_origin, tmp_path):
    raw = wl.text_like(700_001, 41).tobytes()
    gz = wl.gzip_member(raw)
    objs = {
        "/g/text.json.gz": (gz, "application/json", '"gz-ok"'),
        "/g/broken.gz": (gz[: len(gz) // 2], "application/json", '"gz-bad"'),          # truncated member: served as stored (G4)
        "/g/plain.bin": (wl.synth_bytes(90_001, 42).tobytes(), "application/octet-stream", '"id"'),
    }
    d = tmp_path / "cache"
    d.mkdir()
    lines = []
    for path, (data, ctype, etag) in objs.items():
        enc = "gzip" if path.endswith(".gz") else None
        PolicyOrigin.objects[path] = (data, {"Content-Encoding": "gzip"} if enc else {})
        hx = o.md5(o.cache_key("storage.googleapis.com", path)).hex()
        (d / hx).write_bytes(data)
        lines.append(f"{hx} {len(data)} -1 application/octet-stream \"{hashlib.md5(data).hexdigest()}\" {enc or '-'} -")
    (d / "manifest").write_text("\n".join(lines) + "\n")
    gpu_port, cpu_port = free_port(), free_port()
    p = _start_proxy(gpu_port, policy_origin, ["--arena-gib", "0.5", "--gunzip"])
    cpu = subprocess.Popen([os.path.join(ROOT, "oracle", "_build", "nz_cpu_server"), "--root", str(d), "--listen", f"127.0.0.1:{cpu_port}", "--gunzip"],
                           start_new_session=True)
    try:
        wait_port(cpu_port)
        log = []
        for path in objs:
            log += [("GET", path, {}), ("HEAD", path, {}), ("GET", path, {"Accept-Encoding": "gzip"}), ("GET", path, {"Accept-Encoding": "br, GZIP;q=0.5"}),
                    ("GET", path, {"Range": "bytes=100-199"}), ("GET", path, {"Range": "bytes=-9"}), ("GET", path, {"Range": "bytes=699990-"}),
                    ("GET", path, {"Range": "bytes=5-9,300000-300100"}), ("GET", path, {"Range": "bytes=100-199", "Accept-Encoding": "gzip"}),
                    ("GET", path, {"If-None-Match": '"%s"' % hashlib.md5(objs[path][0]).hexdigest()})]
        for method, path, hdrs in log:
            sg, hg, bg = fetch(gpu_port, method, "storage.googleapis.com", path, hdrs)
            sc, hc, bc = fetch(cpu_port, method, "storage.googleapis.com", path, hdrs)
            assert sg == sc, (method, path, hdrs, sg, sc)
            assert normalise(hg, bg) == normalise(hc, bc), (method, path, hdrs, normalise(hg, bg)[0], normalise(hc, bc)[0])
        s, hd, b = fetch(gpu_port, "GET", "storage.googleapis.com", "/g/text.json.gz", {})
        assert s == 200 and b == raw and "content-encoding" not in hd and hd["content-length"] == str(len(raw)) and hd["etag"].startswith('W/')
        s, hd, b = fetch(gpu_port, "GET", "storage.googleapis.com", "/g/broken.gz", {})
        assert s == 200 and b == objs["/g/broken.gz"][0] and hd.get("content-encoding") == "gzip"
        for path in objs:
            assert PolicyOrigin.hits[path] == 1, (path, PolicyOrigin.hits[path])
    finally:
        _stop(p)
        try:
            os.killpg(cpu.pid, signal.SIGTERM)
        except ProcessLookupError:
            pass


def test_private_s3_vhost_signs_fills_with_sigv2():
    """The conf's private-S3 server block (nginx.conf:138-166, commented out in the reference): a fill goes upstream over
    HTTP/1.1 as `Host: <bucket>.s3.amazonaws.com` with `x-amz-date` and `Authorization: AWS <access>:<base64 HMAC-SHA1>`
    over "GET\\n\\n\\n\\nx-amz-date:<date>\\n/<bucket><path>".  The fake S3 below checks the signature the way S3 does."""
    import base64
    import hmac

    bucket, access, secret = "my-datasets", "AKIDEXAMPLE", "wJalrXUtnFEMI/K7MDENG+bPxRfiCYEXAMPLEKEY"
    data = wl.synth_bytes(777_777, 91).tobytes()
    seen = []

    class S3(http.server.BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def do_GET(self):
            date, auth, host = self.headers.get("x-amz-date", ""), self.headers.get("Authorization", ""), self.headers.get("Host", "")
            want = base64.b64encode(hmac.new(secret.encode(), f"GET\n\n\n\nx-amz-date:{date}\n/{bucket}{self.path}".encode(), "sha1").digest()).decode()
            ok = auth == f"AWS {access}:{want}" and host == f"{bucket}.s3.amazonaws.com" and self.path == "/train/shard-0001.tar"
            seen.append((ok, self.request_version, host))
            body = data if ok else b"<Error><Code>SignatureDoesNotMatch</Code></Error>"
            self.send_response(200 if ok else 403)
            self.send_header("Content-Length", str(len(body)))
            self.send_header("Content-Type", "application/x-tar")
            self.send_header("ETag", '"%s"' % hashlib.md5(data).hexdigest())
            self.send_header("Accept-Ranges", "bytes")
            self.send_header("Connection", "close")
            self.end_headers()
            self.wfile.write(body)

        def log_message(self, *a):
            pass

    oport, port = free_port(), free_port()
    httpd = http.server.ThreadingHTTPServer(("127.0.0.1", oport), S3)
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    p = subprocess.Popen([os.path.join(ROOT, "nezha_b200", "_build", "nzproxy"), "--listen", f"127.0.0.1:{port}", "--arena-gib", "0.5", "--max-conns", "8",
                          "--window-mib", "1", "--s3-private", f"{bucket}:{access}:{secret}", "--upstream", f"s3.amazonaws.com=127.0.0.1:{oport}"],
                         start_new_session=True)
    try:
        wait_port(port, 120)
        for _ in range(2):
            s, hd, b = fetch(port, "GET", "s3.amazonaws.com", "/train/shard-0001.tar", {})
            assert s == 200 and b == data and hd.get("content-type") == "application/x-tar"
        s, _, b = fetch(port, "GET", "s3.amazonaws.com", "/train/shard-0001.tar", {"Range": "bytes=700000-700099"})
        assert s == 206 and b == data[700000:700100]
        assert seen == [(True, "HTTP/1.1", f"{bucket}.s3.amazonaws.com")]            # one signed fill, then hits
    finally:
        _stop(p)
        httpd.shutdown()


def test_nzproxy_answers_like_aiohttp_on_the_wire(servers, tmp_path):
    """The GPU front end beside a third-party RFC 7232/7233 implementation (aiohttp's file server), same requests as
    tests/test_third_party_witness.py: first a miss filled from the origin, then hits from HBM."""
    from test_third_party_witness import AGREED, AGREED_AIO, DATA as WDATA, L as WL, MTIME, Aiohttp, same_answer, with_etag

    f = tmp_path / "obj.bin"
    f.write_bytes(WDATA)
    os.utime(f, (MTIME, MTIME))
    aio = Aiohttp(str(tmp_path))
    host, path = "storage.googleapis.com", "/witness/obj.bin"
    OBJECTS[path] = (host, WDATA, "application/octet-stream", aio.etag, None)
    gpu_port, _ = servers
    try:
        for method, hd in AGREED + AGREED_AIO:
            hd = with_etag(hd, aio.etag)
            st, hdrs, body = fetch(gpu_port, method, host, path, hd)
            theirs = aio.ask(method, hd)
            if st == 416:
                assert theirs[0] == 416 and hdrs.get("content-range", "").endswith(f"*/{WL}"), (method, hd, theirs[:3])
                continue
            same_answer((st, hdrs.get("content-range"), hdrs.get("content-length"), hdrs.get("content-type", ""), body), theirs, "aiohttp", (method, hd))
            if st in (200, 206):
                assert hdrs.get("etag") == aio.etag and hdrs.get("last-modified") == LM and hdrs.get("server") == "nginx"
    finally:
        del OBJECTS[path]


def test_request_head_rules_gpu_front_end_vs_cpu_server(servers):
    """SPEC P on the wire, raw heads: duplicate / ignored / invalid header lines, lower-case methods, HTTP/2.0, HTTP/1.1
    without Host, keep-alive on 1.0 - nzproxy and the nginx-restated CPU server must give the same status, Connection,
    Content-Range and body (the CPU server alone is pinned to the expected values in tests/test_cpu_server_wire.py)."""
    from test_cpu_server_wire import P_CASES, raw_exchange

    gpu_port, cpu_port = servers
    host, path = "www.cs.toronto.edu", "/~kriz/cifar-100-python.tar.gz"
    fetch(gpu_port, "GET", host, path, {"Range": "bytes=0-0"})                  # make it a hit
    for head, status, _crange, conn in P_CASES(host, path, OBJECTS[path][3]):
        g = raw_exchange(gpu_port, head, 20)
        c = raw_exchange(cpu_port, head, 20)
        assert g[0] == c[0] == status, (head, g[0], c[0], status)
        assert g[1].get("connection") == c[1].get("connection") == conn, (head, g[1].get("connection"), c[1].get("connection"))
        assert g[1].get("content-range") == c[1].get("content-range"), head
        assert g[2] == c[2], (head, len(g[2]), len(c[2]))
