// nezha_b200/host/webhook.cc — the hostAliases mutating admission webhook (the "pods unchanged" half).
//
// Compiled restatement of /root/reference/app/webhook/webhook.go + pkg/controller/util.go (Go is
// absent from this image).  Same wire contract: POST /mutate-job and /mutate-deployment take an
// admission/v1beta1 AdmissionReview (Content-Type must be exactly application/json, else an empty
// 200 — webhook.go:171-175) and answer
//   {"response":{"uid":<request uid>,"allowed":true,"patch":base64(PATCH),"patchType":"JSONPatch"}}
// with PATCH = `[{"op": "add", "path": "/spec/template/spec/hostAliases", "value": [...] }]`
// (template verbatim from webhook.go:32), existing hostAliases first (webhook.go:92-95).
// Config: YAML list of {name, app, label, hostAliases:[{ip, hostnames:[...]}]}
// (pkg/controller/controller.go:25-30), matched with app==labelKey && label==labelValue
// (util.go:25-33), re-read every 10 s (webhook.go:221-234).
// Consciously fixed vs the reference: labels are visited in sorted key order (Go map order made
// the winner random), a decode failure answers allowed:false instead of panicking (webhook.go:192),
// and the config pointer swap is locked.
#include "webhook.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>

namespace nzw {

// ---------------------------------------------------------------------------
// minimal JSON DOM
// ---------------------------------------------------------------------------
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    std::string s;                                   // Str value or Num literal
    std::vector<Json> a;
    std::vector<std::pair<std::string, Json>> o;     // insertion order; duplicate keys: last wins on lookup
    const Json* get(const std::string& k) const {
        const Json* r = nullptr;
        if (kind == Obj)
            for (const auto& kv : o)
                if (kv.first == k) r = &kv.second;
        return r;
    }
};

namespace {
struct Parser {
    const char* p;
    const char* e;
    [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("invalid JSON: ") + m); }
    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    void put_utf8(std::string& out, uint32_t c) {
        if (c < 0x80) out.push_back((char)c);
        else if (c < 0x800) { out.push_back((char)(0xC0 | (c >> 6))); out.push_back((char)(0x80 | (c & 0x3F))); }
        else if (c < 0x10000) { out.push_back((char)(0xE0 | (c >> 12))); out.push_back((char)(0x80 | ((c >> 6) & 0x3F))); out.push_back((char)(0x80 | (c & 0x3F))); }
        else { out.push_back((char)(0xF0 | (c >> 18))); out.push_back((char)(0x80 | ((c >> 12) & 0x3F))); out.push_back((char)(0x80 | ((c >> 6) & 0x3F))); out.push_back((char)(0x80 | (c & 0x3F))); }
    }
    uint32_t hex4() {
        if (e - p < 4) fail("short \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string str() {
        if (p >= e || *p != '"') fail("expected string");
        ++p;
        std::string out;
        while (p < e && *p != '"') {
            unsigned char c = (unsigned char)*p++;
            if (c < 0x20) fail("control character in string");
            if (c != '\\') { out.push_back((char)c); continue; }
            if (p >= e) fail("short escape");
            const char x = *p++;
            switch (x) {
            case '"': out.push_back('"'); break;
            case '\\': out.push_back('\\'); break;
            case '/': out.push_back('/'); break;
            case 'b': out.push_back('\b'); break;
            case 'f': out.push_back('\f'); break;
            case 'n': out.push_back('\n'); break;
            case 'r': out.push_back('\r'); break;
            case 't': out.push_back('\t'); break;
            case 'u': {
                uint32_t c1 = hex4();
                if (c1 >= 0xD800 && c1 < 0xDC00 && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                    p += 2;
                    const uint32_t c2 = hex4();
                    if (c2 >= 0xDC00 && c2 < 0xE000) c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
                    else { put_utf8(out, 0xFFFD); c1 = c2; }
                }
                put_utf8(out, c1);
                break;
            }
            default: fail("bad escape");
            }
        }
        if (p >= e) fail("unterminated string");
        ++p;
        return out;
    }
    Json value(int depth) {
        if (depth > 200) fail("too deep");
        ws();
        if (p >= e) fail("unexpected end");
        Json j;
        const char c = *p;
        if (c == '{') {
            ++p;
            j.kind = Json::Obj;
            ws();
            if (p < e && *p == '}') { ++p; return j; }
            for (;;) {
                ws();
                std::string k = str();
                ws();
                if (p >= e || *p != ':') fail("expected ':'");
                ++p;
                j.o.emplace_back(std::move(k), value(depth + 1));
                ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            ++p;
            j.kind = Json::Arr;
            ws();
            if (p < e && *p == ']') { ++p; return j; }
            for (;;) {
                j.a.push_back(value(depth + 1));
                ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            j.kind = Json::Str;
            j.s = str();
        } else if (c == 't' && e - p >= 4 && !memcmp(p, "true", 4)) { j.kind = Json::Bool; j.b = true; p += 4; }
        else if (c == 'f' && e - p >= 5 && !memcmp(p, "false", 5)) { j.kind = Json::Bool; j.b = false; p += 5; }
        else if (c == 'n' && e - p >= 4 && !memcmp(p, "null", 4)) { p += 4; }
        else if (c == '-' || (c >= '0' && c <= '9')) {
            const char* s0 = p;
            if (*p == '-') ++p;
            if (p >= e || *p < '0' || *p > '9') fail("bad number");
            while (p < e && ((*p >= '0' && *p <= '9') || *p == '.' || *p == 'e' || *p == 'E' || *p == '+' || *p == '-')) ++p;
            j.kind = Json::Num;
            j.s.assign(s0, p);
        } else {
            fail("unexpected character");
        }
        return j;
    }
};

Json parse_json(const std::string& text) {
    Parser ps{text.data(), text.data() + text.size()};
    Json j = ps.value(0);
    ps.ws();
    if (ps.p != ps.e) ps.fail("trailing data");
    return j;
}

// encoding/json string form: HTML-safe escapes, \u00XX for other control characters
void go_string(std::string& out, const std::string& s) {
    static const char* hexd = "0123456789abcdef";
    out.push_back('"');
    for (size_t i = 0; i < s.size(); ++i) {
        const unsigned char c = (unsigned char)s[i];
        if (c == '"') out += "\\\"";
        else if (c == '\\') out += "\\\\";
        else if (c == '\n') out += "\\n";
        else if (c == '\r') out += "\\r";
        else if (c == '\t') out += "\\t";
        else if (c < 0x20 || c == '<' || c == '>' || c == '&') { out += "\\u00"; out.push_back(hexd[c >> 4]); out.push_back(hexd[c & 15]); }
        else if (c == 0xE2 && i + 2 < s.size() && (unsigned char)s[i + 1] == 0x80 && ((unsigned char)s[i + 2] == 0xA8 || (unsigned char)s[i + 2] == 0xA9)) {
            out += (unsigned char)s[i + 2] == 0xA8 ? "\\u2028" : "\\u2029";
            i += 2;
        } else out.push_back((char)c);
    }
    out.push_back('"');
}

std::string base64(const std::string& in) {
    static const char* tb = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    std::string out;
    size_t i = 0;
    for (; i + 3 <= in.size(); i += 3) {
        const uint32_t v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8) | (unsigned char)in[i + 2];
        out.push_back(tb[v >> 18]); out.push_back(tb[(v >> 12) & 63]); out.push_back(tb[(v >> 6) & 63]); out.push_back(tb[v & 63]);
    }
    if (in.size() - i == 1) {
        const uint32_t v = (unsigned char)in[i] << 16;
        out.push_back(tb[v >> 18]); out.push_back(tb[(v >> 12) & 63]); out += "==";
    } else if (in.size() - i == 2) {
        const uint32_t v = ((unsigned char)in[i] << 16) | ((unsigned char)in[i + 1] << 8);
        out.push_back(tb[v >> 18]); out.push_back(tb[(v >> 12) & 63]); out.push_back(tb[(v >> 6) & 63]); out.push_back('=');
    }
    return out;
}

// ---- YAML subset: block sequences / mappings, plain or quoted scalars, comments ----
struct Line { int indent; std::string text; };

std::string unquote(std::string v) {
    while (!v.empty() && (v.back() == ' ' || v.back() == '\t' || v.back() == '\r')) v.pop_back();
    size_t a = 0;
    while (a < v.size() && (v[a] == ' ' || v[a] == '\t')) ++a;
    v.erase(0, a);
    if (v.size() >= 2 && v.front() == '"' && v.back() == '"') {
        std::string out;
        for (size_t i = 1; i + 1 < v.size(); ++i) {
            if (v[i] == '\\' && i + 2 < v.size()) {
                const char x = v[++i];
                out.push_back(x == 'n' ? '\n' : x == 't' ? '\t' : x);
            } else out.push_back(v[i]);
        }
        return out;
    }
    if (v.size() >= 2 && v.front() == '\'' && v.back() == '\'') {
        std::string out;
        for (size_t i = 1; i + 1 < v.size(); ++i) {
            if (v[i] == '\'' && v[i + 1] == '\'' && i + 2 < v.size()) ++i;
            out.push_back(v[i]);
        }
        return out;
    }
    // plain scalar: cut a trailing comment
    const size_t h = v.find(" #");
    if (h != std::string::npos) { v.erase(h); while (!v.empty() && v.back() == ' ') v.pop_back(); }
    if (v == "~" || v == "null") return "";
    return v;
}

std::vector<Line> yaml_lines(const std::string& text) {
    std::vector<Line> out;
    size_t pos = 0;
    while (pos <= text.size()) {
        size_t nl = text.find('\n', pos);
        if (nl == std::string::npos) nl = text.size();
        std::string ln = text.substr(pos, nl - pos);
        pos = nl + 1;
        if (!ln.empty() && ln.back() == '\r') ln.pop_back();
        size_t ind = 0;
        while (ind < ln.size() && ln[ind] == ' ') ++ind;
        if (ind == ln.size() || ln[ind] == '#') continue;
        if (ln.compare(ind, 3, "---") == 0 && ln.size() == ind + 3) continue;
        out.push_back({(int)ind, ln.substr(ind)});
        if (nl == text.size()) break;
    }
    return out;
}

// splits "key: value" (value may be empty); false when the text is not a mapping entry
bool split_kv(const std::string& t, std::string& k, std::string& v) {
    size_t i = 0;
    if (!t.empty() && (t[0] == '"' || t[0] == '\'')) {
        const char q = t[0];
        i = t.find(q, 1);
        if (i == std::string::npos) return false;
        ++i;
    } else {
        while (i < t.size() && !(t[i] == ':' && (i + 1 == t.size() || t[i + 1] == ' '))) ++i;
    }
    if (i >= t.size() || t[i] != ':') return false;
    k = unquote(t.substr(0, i));
    v = i + 1 < t.size() ? t.substr(i + 1) : "";
    return true;
}
}  // namespace

std::vector<Config> parse_config(const std::string& text) {
    // top level: sequence of mappings; hostAliases: sequence of mappings with ip / hostnames(sequence of scalars)
    const std::vector<Line> L = yaml_lines(text);
    std::vector<Config> out;
    size_t i = 0;
    auto strip_dash = [](const Line& l, int& ind, std::string& rest) -> bool {
        if (l.text.size() >= 1 && l.text[0] == '-' && (l.text.size() == 1 || l.text[1] == ' ')) {
            size_t j = 1;
            while (j < l.text.size() && l.text[j] == ' ') ++j;
            ind = l.indent + (int)j;
            rest = l.text.substr(j);
            return true;
        }
        return false;
    };
    while (i < L.size()) {
        int item_ind;
        std::string rest;
        if (!strip_dash(L[i], item_ind, rest)) throw std::runtime_error("yaml: expected a sequence of configs");
        const int seq_ind = L[i].indent;
        Config cfg;
        // the mapping of this item: first entry on the dash line, the rest at item_ind
        std::vector<std::pair<int, std::string>> ents;
        if (!rest.empty()) ents.emplace_back((int)i, rest);
        size_t j = i + 1;
        while (j < L.size() && L[j].indent > seq_ind) {
            // a '-' line at the key's own indentation is an item of that key's sequence, not a new key
            if (L[j].indent == item_ind && !(L[j].text[0] == '-' && (L[j].text.size() == 1 || L[j].text[1] == ' '))) ents.emplace_back((int)j, L[j].text);
            ++j;
        }
        for (const auto& en : ents) {
            std::string k, v;
            if (!split_kv(en.second, k, v)) throw std::runtime_error("yaml: expected key: value");
            const std::string val = unquote(v);
            if (k == "name") cfg.name = val;
            else if (k == "app") cfg.app = val;
            else if (k == "label") cfg.label = val;
            else if (k == "hostAliases") {
                // nested sequence: lines after this entry with indent > item_ind (or == item_ind for "- " style)
                size_t a = (size_t)en.first + 1;
                while (a < j) {
                    int al_ind;
                    std::string arest;
                    if (L[a].indent < item_ind || !strip_dash(L[a], al_ind, arest)) break;
                    const int aseq = L[a].indent;
                    HostAlias ha;
                    std::vector<std::pair<size_t, std::string>> aent;
                    if (!arest.empty()) aent.emplace_back(a, arest);
                    size_t b = a + 1;
                    while (b < j && L[b].indent > aseq) {
                        if (L[b].indent == al_ind && !(L[b].text[0] == '-' && (L[b].text.size() == 1 || L[b].text[1] == ' '))) aent.emplace_back(b, L[b].text);
                        ++b;
                    }
                    for (const auto& ae : aent) {
                        std::string ak, av;
                        if (!split_kv(ae.second, ak, av)) throw std::runtime_error("yaml: expected key: value in hostAliases");
                        if (ak == "ip") ha.ip = unquote(av);
                        else if (ak == "hostnames" && unquote(av).size() >= 2 && unquote(av).front() == '[') {
                            // one-line flow sequence: hostnames: [a, "b"]
                            std::string f = unquote(av);
                            f = f.substr(1, f.rfind(']') == std::string::npos ? std::string::npos : f.rfind(']') - 1);
                            size_t q = 0;
                            while (q <= f.size()) {
                                size_t c = f.find(',', q);
                                if (c == std::string::npos) c = f.size();
                                const std::string item = unquote(f.substr(q, c - q));
                                if (!item.empty()) ha.hostnames.push_back(item);
                                q = c + 1;
                            }
                        } else if (ak == "hostnames") {
                            for (size_t h = ae.first + 1; h < b; ++h) {
                                int hi;
                                std::string hrest;
                                if (L[h].indent < al_ind || !strip_dash(L[h], hi, hrest)) break;
                                ha.hostnames.push_back(unquote(hrest));
                            }
                        }
                    }
                    cfg.aliases.push_back(std::move(ha));
                    a = b;
                }
            }
        }
        out.push_back(std::move(cfg));
        i = j;
    }
    return out;
}

static const std::vector<HostAlias>* aliases_by_kv(const std::string& k, const std::string& v, const std::vector<Config>& cfg) {
    for (const Config& c : cfg)                       // util.go:25-33
        if (c.app == k && c.label == v) return &c.aliases;
    return nullptr;
}

static std::string marshal_aliases(const std::vector<HostAlias>& v) {
    std::string s = "[";
    for (size_t i = 0; i < v.size(); ++i) {
        if (i) s.push_back(',');
        s.push_back('{');
        bool first = true;
        if (!v[i].ip.empty()) { s += "\"ip\":"; go_string(s, v[i].ip); first = false; }
        if (!v[i].hostnames.empty()) {
            if (!first) s.push_back(',');
            s += "\"hostnames\":[";
            for (size_t h = 0; h < v[i].hostnames.size(); ++h) {
                if (h) s.push_back(',');
                go_string(s, v[i].hostnames[h]);
            }
            s.push_back(']');
        }
        s.push_back('}');
    }
    s.push_back(']');
    return s;
}

static std::string error_response(const std::string& uid, const std::string& msg) {
    std::string s = "{\"response\":{\"uid\":";
    go_string(s, uid);
    s += ",\"allowed\":false,\"status\":{\"metadata\":{},\"message\":";
    go_string(s, msg);
    s += "}}}";
    return s;
}

std::string serve(const std::vector<Config>& cfg, const std::string& path, const char* content_type, const std::string& body) {
    if (!content_type || std::string(content_type) != "application/json") return "";      // webhook.go:171-175
    const bool job = path == "/mutate-job";
    std::string uid;
    try {
        const Json ar = parse_json(body);
        if (ar.kind != Json::Obj) throw std::runtime_error("invalid JSON: not an object");
        const Json* req = ar.get("request");
        const Json* ju = req ? req->get("uid") : nullptr;
        if (ju && ju->kind == Json::Str) uid = ju->s;
        const Json* res = req ? req->get("resource") : nullptr;
        auto field = [&](const char* k) { const Json* f = res ? res->get(k) : nullptr; return f && f->kind == Json::Str ? f->s : std::string(); };
        const bool match = job ? (field("group") == "batch" && field("version") == "v1" && field("resource") == "jobs")
                               : (field("group") == "extensions" && field("version") == "v1beta1" && field("resource") == "deployments");
        if (!match) return "{}";                                                               // webhook.go:73-76,190-193
        const Json* obj = req->get("object");
        if (!obj || obj->kind != Json::Obj) throw std::runtime_error("object is not a JSON object");
        std::map<std::string, std::string> labels;
        if (const Json* md = obj->get("metadata"); md && md->kind == Json::Obj)
            if (const Json* lb = md->get("labels"); lb && lb->kind != Json::Null) {
                if (lb->kind != Json::Obj) throw std::runtime_error("labels must be a string map");
                for (const auto& kv : lb->o) {
                    if (kv.second.kind != Json::Str) throw std::runtime_error("labels must be a string map");
                    labels[kv.first] = kv.second.s;
                }
            }
        std::vector<HostAlias> existing;
        const Json* sp = obj->get("spec");
        const Json* tp = sp && sp->kind == Json::Obj ? sp->get("template") : nullptr;
        const Json* ps = tp && tp->kind == Json::Obj ? tp->get("spec") : nullptr;
        if (const Json* ha = ps && ps->kind == Json::Obj ? ps->get("hostAliases") : nullptr; ha && ha->kind != Json::Null) {
            if (ha->kind != Json::Arr) throw std::runtime_error("bad hostAliases types");
            for (const Json& a : ha->a) {
                if (a.kind != Json::Obj) throw std::runtime_error("hostAliases entry is not an object");
                HostAlias h;
                if (const Json* ip = a.get("ip"); ip && ip->kind != Json::Null) {
                    if (ip->kind != Json::Str) throw std::runtime_error("bad hostAliases types");
                    h.ip = ip->s;
                }
                if (const Json* hn = a.get("hostnames"); hn && hn->kind != Json::Null) {
                    if (hn->kind != Json::Arr) throw std::runtime_error("bad hostAliases types");
                    for (const Json& x : hn->a) {
                        if (x.kind != Json::Str) throw std::runtime_error("bad hostAliases types");
                        h.hostnames.push_back(x.s);
                    }
                }
                existing.push_back(std::move(h));
            }
        }
        std::string patch;
        for (const auto& kv : labels) {                                                        // sorted key order
            const std::vector<HostAlias>* al = aliases_by_kv(kv.first, kv.second, cfg);
            if (al && !al->empty()) {
                std::vector<HostAlias> merged = existing;                                      // existing first (webhook.go:92-95)
                merged.insert(merged.end(), al->begin(), al->end());
                patch = "[{\"op\": \"add\", \"path\": \"/spec/template/spec/hostAliases\", \"value\": " + marshal_aliases(merged) + " }]";
            }
        }
        std::string s = "{\"response\":{\"uid\":";
        go_string(s, uid);
        s += ",\"allowed\":true";
        if (!patch.empty()) s += ",\"patch\":\"" + base64(patch) + "\",\"patchType\":\"JSONPatch\"";
        s += "}}";
        return s;
    } catch (const std::exception& e) {
        return error_response(uid, e.what());
    }
}

}  // namespace nzw

extern "C" int nzw_serve(const char* config_yaml, const char* path, const char* content_type, const char* body, size_t body_len,
                         char* out, size_t cap) {
    try {
        const std::vector<nzw::Config> cfg = nzw::parse_config(config_yaml ? config_yaml : "");
        const std::string r = nzw::serve(cfg, path ? path : "", content_type, std::string(body ? body : "", body_len));
        if (r.size() + 1 > cap) return -1;
        memcpy(out, r.c_str(), r.size() + 1);
        return (int)r.size();
    } catch (const std::exception&) {
        return -2;   // unparseable config
    }
}
