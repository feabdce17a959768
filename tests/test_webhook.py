"""-m "not gpu": the hostAliases mutating webhook.

Product = nezha_b200/host/webhook.cc (+ nzwebhook HTTPS server); checker =
oracle/webhook_oracle.py, a line-by-line restatement of the reference's
app/webhook/webhook.go.  Fixtures are the reference's own example objects
(examples/webhook-test.yaml, examples/demo/kubeflow-test.yaml) and config shapes
(deploy/mutatingwebhook.yaml:14-25, examples/demo/setup.sh:21-38, examples/configmap.yaml),
restated inline because /root/reference does not exist on the GPU box.
"""
import base64
import ctypes as C
import json
import os
import socket
import ssl
import subprocess
import time

import pytest
from hypothesis import given, settings, strategies as st

from oracle import webhook_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = C.CDLL(os.path.join(ROOT, "nezha_b200", "_build", "libnezha_host.so"))
LIB.nzw_serve.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]


def product(cfg_yaml: str, path: str, ctype, body: bytes):
    out = C.create_string_buffer(1 << 20)
    n = LIB.nzw_serve(cfg_yaml.encode(), path.encode(), ctype.encode() if ctype is not None else None, body, len(body), out, len(out))
    assert n >= 0, n
    return out.raw[:n]


# examples/demo/setup.sh:21-38 with the ClusterIP filled in
CFG_DEMO = """
- name: dataset
  app: app.kubernetes.io/deploy-manager
  label: ksonnet
  hostAliases:
  - ip: "10.96.12.34"
    hostnames:
      - "download.tensorflow.org"
      - "yann.lecun.com"
      - "www.cs.toronto.edu"
      - "storage.googleapis.com"
      - "images.cocodataset.org"
"""
# deploy/mutatingwebhook.yaml:14-25 shape (no `app` key) and examples/configmap.yaml shape (deeper indent, two aliases)
CFG_DEFAULT = """
- name: dataset
  label: dataset
  hostAliases:
  - ip: "1.2.3.4"
    hostnames:
    - "example.com"
"""
CFG_TWO = """
      - name: application
        app: app
        label: gcs          # examples/webhook-test.yaml labels app: gcs
        hostAliases:
            - ip: "127.0.0.1"
              hostnames:
                - "foo.local"
                - "bar.local"
            - ip: "10.1.2.3"
              hostnames:
                - "foo.remote"
                - "bar.remote"
      - name: second
        app: tier
        label: 'train'
        hostAliases:
            - ip: 10.9.9.9
              hostnames: ["ignored-flow-style"]
"""


def review(kind, labels, host_aliases=None, uid="705ab4f5-6393-11e8-b7cc-42010a800002", gvr=None):
    if kind == "job":      # examples/demo/kubeflow-test.yaml:8-22
        obj = {"apiVersion": "batch/v1", "kind": "Job", "metadata": {"name": "nezha-job-test", "namespace": "nezha-test", "labels": labels},
               "spec": {"template": {"spec": {"restartPolicy": "OnFailure", "containers": [{"name": "wget", "image": "mwendler/wget",
                        "args": ["http://www.cs.toronto.edu/~kriz/cifar-100-python.tar.gz"]}]}}}}
        g = {"group": "batch", "version": "v1", "resource": "jobs"}
    else:                  # examples/webhook-test.yaml:8-28
        obj = {"apiVersion": "extensions/v1beta1", "kind": "Deployment", "metadata": {"name": "webhook-dp-test", "namespace": "test", "labels": labels},
               "spec": {"replicas": 1, "template": {"metadata": {"labels": labels}, "spec": {"containers": [{"name": "sleep", "image": "centos"}]}}}}
        g = {"group": "extensions", "version": "v1beta1", "resource": "deployments"}
    if host_aliases is not None:
        obj["spec"]["template"]["spec"]["hostAliases"] = host_aliases
    return {"kind": "AdmissionReview", "apiVersion": "admission.k8s.io/v1beta1",
            "request": {"uid": uid, "kind": {"group": g["group"], "version": g["version"], "kind": obj["kind"]}, "resource": gvr or g,
                        "namespace": obj["metadata"]["namespace"], "operation": "CREATE", "object": obj}}


def both(cfg, path, ar, ctype="application/json"):
    body = json.dumps(ar).encode() if not isinstance(ar, bytes) else ar
    got = product(cfg, path, ctype, body)
    want = wo.serve(path, ctype, body, wo.parse_config(cfg))
    return got, want


def test_demo_job_gets_the_patch_template_verbatim():
    got, want = both(CFG_DEMO, "/mutate-job", review("job", {"app.kubernetes.io/deploy-manager": "ksonnet"}))
    assert json.loads(got) == json.loads(want)
    r = json.loads(got)["response"]
    assert r["allowed"] is True and r["patchType"] == "JSONPatch" and r["uid"] == "705ab4f5-6393-11e8-b7cc-42010a800002"
    patch = base64.b64decode(r["patch"]).decode()
    # webhook.go:32 template, byte for byte, around Go's compact JSON of []HostAlias
    assert patch == ('[{"op": "add", "path": "/spec/template/spec/hostAliases", "value": [{"ip":"10.96.12.34","hostnames":'
                     '["download.tensorflow.org","yann.lecun.com","www.cs.toronto.edu","storage.googleapis.com","images.cocodataset.org"]}] }]')
    assert json.loads(patch)[0]["op"] == "add"


def test_deployment_existing_aliases_come_first():
    existing = [{"ip": "9.9.9.9", "hostnames": ["old.local"], "unknownField": 1}]
    got, want = both(CFG_TWO, "/mutate-deployment", review("deployment", {"app": "gcs"}, existing))
    assert json.loads(got) == json.loads(want)
    val = json.loads(base64.b64decode(json.loads(got)["response"]["patch"]))[0]["value"]
    assert val == [{"ip": "9.9.9.9", "hostnames": ["old.local"]}, {"ip": "127.0.0.1", "hostnames": ["foo.local", "bar.local"]},
                   {"ip": "10.1.2.3", "hostnames": ["foo.remote", "bar.remote"]}]


def test_no_match_wrong_gvr_bad_content_type_bad_json():
    got, want = both(CFG_DEMO, "/mutate-job", review("job", {"app": "other"}))
    assert json.loads(got) == json.loads(want) == {"response": {"uid": "705ab4f5-6393-11e8-b7cc-42010a800002", "allowed": True}}
    # deploy/mutatingwebhook.yaml default config has no `app`: key "" never equals a label key
    got, want = both(CFG_DEFAULT, "/mutate-job", review("job", {"dataset": "dataset", "": "dataset"}))
    assert json.loads(got) == json.loads(want)
    # a Job sent to the deployment endpoint: response omitted (webhook.go:73-76,190-193)
    got, want = both(CFG_DEMO, "/mutate-deployment", review("job", {"app.kubernetes.io/deploy-manager": "ksonnet"}))
    assert got == want == b"{}"
    got, want = both(CFG_DEMO, "/mutate-job", review("job", {}), ctype="application/json; charset=utf-8")
    assert got == want == b""                                   # exact match only (webhook.go:171-175)
    got, want = both(CFG_DEMO, "/mutate-job", review("job", {}), ctype=None)
    assert got == want == b""
    for bad in (b"{not json", b"[]", b'{"request":{"resource":{"group":"batch","version":"v1","resource":"jobs"},"object":5}}',
                json.dumps(review("job", {"a": 1})).encode()):
        g = json.loads(product(CFG_DEMO, "/mutate-job", "application/json", bad))
        w = json.loads(wo.serve("/mutate-job", "application/json", bad, wo.parse_config(CFG_DEMO)))
        assert g["response"]["allowed"] is False and w["response"]["allowed"] is False
        assert g["response"]["status"]["message"] and g["response"]["status"]["metadata"] == {}


def test_multiple_matching_labels_deterministic_last_sorted_key_wins():
    cfg = CFG_TWO
    got, want = both(cfg, "/mutate-job", review("job", {"tier": "train", "app": "gcs"}))
    assert json.loads(got) == json.loads(want)
    val = json.loads(base64.b64decode(json.loads(got)["response"]["patch"]))[0]["value"]
    assert val[0]["ip"] == "10.9.9.9"        # "tier" > "app": visited last, wins (documented determinisation)


label_txt = st.text(alphabet="abcdefgz/.-_<>&\"\\ é ", min_size=0, max_size=12)


@settings(max_examples=200, deadline=None)
@given(st.dictionaries(label_txt, label_txt, max_size=4), st.lists(st.fixed_dictionaries({"ip": label_txt, "hostnames": st.lists(label_txt, max_size=3)}), max_size=3),
       st.sampled_from(["job", "deployment"]), label_txt)
def test_differential_generated(labels, existing, kind, uid):
    labels = dict(labels)
    labels.setdefault("app", "gcs")
    got, want = both(CFG_TWO, "/mutate-job" if kind == "job" else "/mutate-deployment", review(kind, labels, existing or None, uid=uid))
    assert json.loads(got) == json.loads(want)
    r = json.loads(got)["response"]
    if "patch" in r:
        assert base64.b64decode(r["patch"]) == base64.b64decode(json.loads(want)["response"]["patch"])   # byte-exact patch


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_https_server_end_to_end_and_config_reload(tmp_path):
    cert, key, cfgf = tmp_path / "tls.crt", tmp_path / "tls.key", tmp_path / "config.yaml"
    subprocess.check_call(["openssl", "req", "-x509", "-newkey", "rsa:2048", "-nodes", "-keyout", str(key), "-out", str(cert), "-days", "2",
                           "-subj", "/CN=hostaliases-injector-webhook-svc.default.svc"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cfgf.write_text(CFG_DEMO)
    port = free_port()
    p = subprocess.Popen([os.path.join(ROOT, "nezha_b200", "_build", "nzwebhook"), "-config-file", str(cfgf), "-tls-cert-file", str(cert),
                          "-tls-private-key-file", str(key), "-listen", f"127.0.0.1:{port}", "-v=5"], stderr=subprocess.DEVNULL)
    try:
        import http.client
        ctx = ssl.create_default_context()
        ctx.check_hostname = False
        ctx.verify_mode = ssl.CERT_NONE
        for _ in range(50):
            try:
                socket.create_connection(("127.0.0.1", port), 0.2).close()
                break
            except OSError:
                time.sleep(0.1)

        def post(path, body, ctype="application/json"):
            c = http.client.HTTPSConnection("127.0.0.1", port, context=ctx, timeout=10)
            c.request("POST", path, body=body, headers={"Content-Type": ctype})
            r = c.getresponse()
            out = r.status, r.read()
            c.close()
            return out

        ar = json.dumps(review("job", {"app.kubernetes.io/deploy-manager": "ksonnet"})).encode()
        st_, body = post("/mutate-job", ar)
        assert st_ == 200 and json.loads(body) == json.loads(wo.serve("/mutate-job", "application/json", ar, wo.parse_config(CFG_DEMO)))
        assert post("/mutate-job", ar, "text/plain") == (200, b"")
        assert post("/nope", ar)[0] == 404
    finally:
        p.kill()
