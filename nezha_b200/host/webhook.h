// nezha_b200/host/webhook.h — hostAliases mutating admission webhook (see webhook.cc)
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace nzw {
struct HostAlias { std::string ip; std::vector<std::string> hostnames; };             // core/v1.HostAlias
struct Config { std::string name, app, label; std::vector<HostAlias> aliases; };     // pkg/controller/controller.go:25-30
std::vector<Config> parse_config(const std::string& yaml_text);                       // throws std::runtime_error
// body of the HTTP response for POST <path>; "" when the content type is not application/json
std::string serve(const std::vector<Config>& cfg, const std::string& path, const char* content_type, const std::string& body);
}  // namespace nzw

extern "C" int nzw_serve(const char* config_yaml, const char* path, const char* content_type, const char* body, size_t body_len,
                         char* out, size_t cap);
