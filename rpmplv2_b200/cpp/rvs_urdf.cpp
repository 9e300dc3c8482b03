// rvs_urdf.cpp — see rvs_urdf.h.
#include "rvs_urdf.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <vector>

namespace rvs {
namespace {

struct JointRec {
    std::string type, parent, child;
    double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0}, axis[3] = {1, 0, 0};     // URDF defaults
    double lower = 0, upper = 0;
};

// value of attribute `name` inside the tag text [tag_begin, tag_end)
bool attribute(const std::string &s, size_t tag_begin, size_t tag_end, const char *name, std::string *value) {
    const std::string key = std::string(name) + "=";
    size_t p = tag_begin;
    while ((p = s.find(key, p)) != std::string::npos && p < tag_end) {
        const char before = p > 0 ? s[p - 1] : ' ';
        if (before == ' ' || before == '\t' || before == '\n' || before == '\r') {
            const size_t q0 = p + key.size();
            if (q0 >= tag_end || (s[q0] != '"' && s[q0] != '\'')) return false;
            const size_t q1 = s.find(s[q0], q0 + 1);
            if (q1 == std::string::npos || q1 > tag_end) return false;
            *value = s.substr(q0 + 1, q1 - q0 - 1);
            return true;
        }
        p += key.size();
    }
    return false;
}

// first `<name ...>` tag inside [from, to); returns its extent
bool find_tag(const std::string &s, size_t from, size_t to, const char *name, size_t *b, size_t *e) {
    const std::string open = std::string("<") + name;
    size_t p = from;
    while ((p = s.find(open, p)) != std::string::npos && p < to) {
        const char after = s[p + open.size()];
        if (after == ' ' || after == '\t' || after == '\n' || after == '\r' || after == '>' || after == '/') {
            const size_t close = s.find('>', p);
            if (close == std::string::npos || close > to) return false;
            *b = p; *e = close;
            return true;
        }
        p += open.size();
    }
    return false;
}

bool triple(const std::string &text, double out[3]) {
    std::istringstream is(text);
    return bool(is >> out[0] >> out[1] >> out[2]);
}

std::string strip_comments(const std::string &in) {
    std::string s;
    size_t p = 0;
    for (;;) {
        const size_t c = in.find("<!--", p);
        if (c == std::string::npos) { s.append(in, p, std::string::npos); break; }
        s.append(in, p, c - p);
        const size_t e = in.find("-->", c);
        if (e == std::string::npos) break;
        p = e + 3;
    }
    return s;
}

}  // namespace

bool chain_from_urdf_text(const std::string &raw, const std::string &robot_type, rvs_robot_desc *out, std::string *error) {
    auto fail = [&](const std::string &m) { if (error) *error = m; return false; };
    if (!out) return fail("null output descriptor");
    const std::string s = strip_comments(raw);
    std::map<std::string, JointRec> by_child;
    size_t p = 0, b, e;
    while (find_tag(s, p, s.size(), "joint", &b, &e)) {
        p = e + 1;
        if (s[e - 1] == '/') continue;                                   // <joint name="..."/> inside <transmission>
        const size_t end = s.find("</joint>", e);
        if (end == std::string::npos) return fail("unterminated <joint>");
        JointRec j;
        std::string v;
        if (!attribute(s, b, e, "type", &j.type)) { p = end; continue; }
        size_t tb, te;
        if (!find_tag(s, e, end, "parent", &tb, &te) || !attribute(s, tb, te, "link", &j.parent)) { p = end; continue; }
        if (!find_tag(s, e, end, "child", &tb, &te) || !attribute(s, tb, te, "link", &j.child)) { p = end; continue; }
        if (find_tag(s, e, end, "origin", &tb, &te)) {
            if (attribute(s, tb, te, "xyz", &v) && !triple(v, j.xyz)) return fail("bad origin xyz");
            if (attribute(s, tb, te, "rpy", &v) && !triple(v, j.rpy)) return fail("bad origin rpy");
        }
        if (find_tag(s, e, end, "axis", &tb, &te) && attribute(s, tb, te, "xyz", &v) && !triple(v, j.axis)) return fail("bad axis");
        if (find_tag(s, e, end, "limit", &tb, &te)) {
            if (attribute(s, tb, te, "lower", &v)) j.lower = std::strtod(v.c_str(), nullptr);
            if (attribute(s, tb, te, "upper", &v)) j.upper = std::strtod(v.c_str(), nullptr);
        }
        by_child[j.child] = j;
        p = end;
    }
    const bool is_xarm6 = robot_type == "xarm6";
    const std::string base = is_xarm6 ? "link_base" : "base_link", tip = is_xarm6 ? "link_eef" : "tool";
    std::vector<JointRec> chain;
    for (std::string link = tip; link != base;) {
        auto it = by_child.find(link);
        if (it == by_child.end()) return fail("link '" + link + "' is not reachable from '" + base + "'");
        chain.push_back(it->second);
        link = it->second.parent;
        if (chain.size() > size_t(RVS_MAX_DOFS) + 1) return fail("chain longer than RVS_MAX_DOFS");
    }
    if (chain.size() < 2) return fail("chain has no moving joint");
    std::memset(out, 0, sizeof(*out));
    const int n = int(chain.size()) - 1;
    for (int k = 0; k <= n; k++) {
        const JointRec &j = chain[chain.size() - 1 - size_t(k)];
        const bool moving = j.type == "revolute" || j.type == "continuous";
        if ((k < n) != moving || (k == n && j.type != "fixed")) return fail("expected n revolute joints followed by one fixed tip joint");
        for (int c = 0; c < 3; c++) { out->joint_xyz[k][c] = j.xyz[c]; out->joint_rpy[k][c] = j.rpy[c]; }
        if (k < n) {
            const double len = std::sqrt(j.axis[0] * j.axis[0] + j.axis[1] * j.axis[1] + j.axis[2] * j.axis[2]);
            if (!(len > 0)) return fail("zero joint axis");
            for (int c = 0; c < 3; c++) out->joint_axis[k][c] = j.axis[c] / len;
            out->limit_lo[k] = float(j.lower);                            // `float lower = ...->limits->lower` (xArm6.cpp:30-35)
            out->limit_hi[k] = float(j.upper);
        }
    }
    out->num_dofs = n;
    out->kind = is_xarm6 ? RVS_ROBOT_XARM6 : (robot_type.rfind("planar", 0) == 0 ? RVS_ROBOT_PLANAR : RVS_ROBOT_SPATIAL);
    return true;
}

bool chain_from_urdf_file(const std::string &path, const std::string &robot_type, rvs_robot_desc *out, std::string *error) {
    std::ifstream f(path);
    if (!f) { if (error) *error = "cannot open " + path; return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    return chain_from_urdf_text(ss.str(), robot_type, out, error);
}

}  // namespace rvs
