"""oracle/webhook_oracle.py — restatement of the reference's mutating admission webhook.

TEST INFRASTRUCTURE ONLY.  Follows /root/reference/app/webhook/webhook.go and
/root/reference/pkg/controller/util.go line by line (Go is absent from the image, so
the reference itself cannot run here: PARITY UNPINNED except for the literals the
reference fixes — the patch template webhook.go:32, the config schema
controller.go:25-30, the fixtures under examples/ and deploy/).

One deliberate determinisation: webhook.go:89-106 ranges over a Go map of labels, so
when several labels match the LAST one visited wins and the visit order is random.
Here (and in the product) labels are visited in sorted key order.
"""
from __future__ import annotations

import base64
import json
from typing import Any, Dict, List, Optional

import yaml

PATCH_TEMPLATE = '[{"op": "add", "path": "/spec/template/spec/hostAliases", "value": %s }]'   # webhook.go:32

DEPLOYMENT_GVR = {"group": "extensions", "version": "v1beta1", "resource": "deployments"}   # webhook.go:72
JOB_GVR = {"group": "batch", "version": "v1", "resource": "jobs"}                          # webhook.go:113


def parse_config(text: str) -> List[Dict[str, Any]]:
    """pkg/controller/util.go:45-58 FileToConfig: yaml.Unmarshal into []Config
    (controller.go:25-30: name, app, label, hostAliases; HostAlias fields fall back to the
    lower-cased Go field names ip / hostnames — vendor/gopkg.in/yaml.v2/yaml.go:392-396)."""
    raw = yaml.safe_load(text) or []
    out = []
    for c in raw:
        c = c or {}
        aliases = []
        for a in c.get("hostAliases") or []:
            a = a or {}
            aliases.append({"ip": str(a.get("ip", "") or ""), "hostnames": [str(h) for h in (a.get("hostnames") or [])]})
        out.append({"name": str(c.get("name", "") or ""), "app": str(c.get("app", "") or ""), "label": str(c.get("label", "") or ""),
                    "aliases": aliases})
    return out


def get_aliases_by_kv(k: str, v: str, config) -> Optional[list]:
    """util.go:25-33"""
    for conf in config:
        if conf["app"] == k and conf["label"] == v:
            return conf["aliases"]
    return None


def go_json(v) -> str:
    """encoding/json.Marshal of []core/v1.HostAlias: compact, HTML-escaped, omitempty on both fields."""
    def alias(a):
        d = {}
        if a.get("ip"):
            d["ip"] = a["ip"]
        if a.get("hostnames"):
            d["hostnames"] = list(a["hostnames"])
        return d
    s = json.dumps([alias(a) for a in v], separators=(",", ":"), ensure_ascii=False)
    return s.replace("<", "\\u003c").replace(">", "\\u003e").replace("&", "\\u0026").replace("\u2028", "\\u2028").replace("\u2029", "\\u2029")


def _typed_aliases(raw) -> list:
    """What survives decoding spec.template.spec.hostAliases into []HostAlias (unknown fields dropped)."""
    out = []
    for a in raw or []:
        if not isinstance(a, dict):
            raise ValueError("hostAliases entry is not an object")
        ip = a.get("ip", "")
        hn = a.get("hostnames") or []
        if not isinstance(ip, str) or not isinstance(hn, list) or any(not isinstance(h, str) for h in hn):
            raise ValueError("bad hostAliases types")
        out.append({"ip": ip, "hostnames": hn})
    return out


def mutate(ar: dict, config, gvr: dict) -> Optional[dict]:
    """mutateDeployments / mutateJobs (webhook.go:70-150).  Returns the AdmissionResponse dict or None."""
    req = ar.get("request") or {}
    res = req.get("resource") or {}
    if {k: res.get(k, "") for k in ("group", "version", "resource")} != gvr:
        return None                                                   # webhook.go:73-76 / 114-117
    obj = req.get("object")
    try:
        if not isinstance(obj, dict):
            raise ValueError("object is not a JSON object")
        meta = obj.get("metadata") or {}
        labels = meta.get("labels") or {}
        if not isinstance(labels, dict) or any(not isinstance(v, str) for v in labels.values()):
            raise ValueError("labels must be a string map")
        spec = (((obj.get("spec") or {}).get("template") or {}).get("spec") or {})
        existing = _typed_aliases(spec.get("hostAliases"))
    except (ValueError, AttributeError) as e:
        return {"uid": "", "allowed": False, "status": {"metadata": {}, "message": str(e)}}      # toAdmissionResponse
    resp: Dict[str, Any] = {"uid": "", "allowed": True}
    for k in sorted(labels):                                          # determinised map order
        aliases = get_aliases_by_kv(k, labels[k], config)
        if aliases:
            merged = existing + aliases if existing else aliases       # webhook.go:92-95: existing first
            patch = PATCH_TEMPLATE % go_json(merged)
            resp["patch"] = base64.b64encode(patch.encode()).decode()
            resp["patchType"] = "JSONPatch"
    return resp


def serve(path: str, content_type: Optional[str], body: bytes, config) -> bytes:
    """serve() (webhook.go:162-205): returns the response body bytes ('' when the content type is wrong)."""
    if content_type != "application/json":
        return b""                                                    # webhook.go:171-175
    gvr = JOB_GVR if path == "/mutate-job" else DEPLOYMENT_GVR
    uid = ""
    try:
        ar = json.loads(body)
        if not isinstance(ar, dict):
            raise ValueError("not an object")
        uid = ((ar.get("request") or {}).get("uid")) or ""
        resp = mutate(ar, config, gvr)
    except (ValueError, AttributeError) as e:
        resp = {"uid": "", "allowed": False, "status": {"metadata": {}, "message": str(e)}}
    out: Dict[str, Any] = {}
    if resp is not None:
        resp["uid"] = uid                                             # webhook.go:192
        out["response"] = resp
    return json.dumps(out, separators=(",", ":")).encode()
