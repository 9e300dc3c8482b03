// oracle/env_shim/yaml-cpp — stand-in for yaml-cpp: every node is UNDEFINED, so env::Environment's YAML constructor runs into its own
// "Workspace center point is not defined!" exception (caught inside the constructor) and returns; the harness then sets the members
// the motion laws read.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <string>
namespace YAML {
class Node {
public:
    Node operator[](const char *) const { return Node(); }
    Node operator[](const std::string &) const { return Node(); }
    Node operator[](std::size_t) const { return Node(); }
    Node operator[](int) const { return Node(); }
    bool IsDefined() const { return false; }
    std::size_t size() const { return 0; }
    template <class T> T as() const { return T(); }
};
inline Node LoadFile(const std::string &) { return Node(); }
}  // namespace YAML
