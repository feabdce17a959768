"""CPU: the manifests in deploy/ start the binaries with arguments those binaries accept.
On a box without a GPU nzproxy must get past argument parsing and then fail loudly in nz_init
(exit 1, "no CPU fallback"); a usage error (exit 2) means the manifest and the parser disagree."""
import os
import signal
import subprocess

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def containers():
    out = {}
    with open(os.path.join(ROOT, "deploy", "nezha-b200.yaml")) as f:
        for doc in yaml.safe_load_all(f):
            if not doc or doc.get("kind") != "Deployment":
                continue
            for c in doc["spec"]["template"]["spec"]["containers"]:
                out[c["name"]] = c
    return out


def test_nzproxy_accepts_the_manifest_arguments(tmp_path):
    c = containers()["nzproxy"]
    args = []
    for a in c["args"]:
        a = a.replace("/var/cache/nezha", str(tmp_path / "pv")).replace("/var/log/nginx/access.log", str(tmp_path / "access.log"))
        a = a.replace("0.0.0.0:80", "127.0.0.1:0")
        args.append(a)
    os.makedirs(tmp_path / "pv", exist_ok=True)
    exe = os.path.join(ROOT, "nezha_b200", "_build", "nzproxy")
    p = subprocess.Popen([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, start_new_session=True)
    try:
        _, err = p.communicate(timeout=20)
    except subprocess.TimeoutExpired:            # a GPU is present: it parsed its arguments and is serving
        os.killpg(p.pid, signal.SIGKILL)
        p.wait()
        return
    assert p.returncode == 1, (p.returncode, err.decode()[-400:])
    assert b"nz_init" in err and b"no CPU fallback" in err


def test_service_and_webhook_objects_keep_the_reference_names():
    """examples/demo/nginx.yaml: Service proxy-cache port 80; deploy/mutatingwebhook.yaml paths."""
    with open(os.path.join(ROOT, "deploy", "nezha-b200.yaml")) as f:
        docs = [d for d in yaml.safe_load_all(f) if d]
    svc = [d for d in docs if d["kind"] == "Service" and d["metadata"]["name"] == "proxy-cache"]
    assert svc and svc[0]["spec"]["ports"][0]["port"] == 80
    hooks = [d for d in docs if d["kind"] == "MutatingWebhookConfiguration"]
    if hooks:
        paths = {w["clientConfig"]["service"]["path"] for w in hooks[0]["webhooks"]}
        assert paths <= {"/mutate-job", "/mutate-deployment"} and paths
