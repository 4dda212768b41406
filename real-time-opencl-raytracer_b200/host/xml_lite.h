// xml_lite.h -- a minimal in-memory XML reader, just enough for the COLLADA dialect the loader
// accepts (elements, attributes, character data, comments, <?...?> and <!...> skipped).
//
// The reference parses XML with pugixml, which it neither vendors nor pins (reference README.md:4,
// ColladaLoader.h:12); this replaces that dependency. The query surface mimics the handful of
// pugixml calls the reference makes: child(name), next_sibling(name), attribute(name), text().
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace xml_lite {

struct Node {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::string text;  // first character-data run directly inside this element (pugixml text())
    bool text_set = false;
    Node* parent = nullptr;
    size_t index_in_parent = 0;
    std::vector<std::unique_ptr<Node>> children;

    const Node* child(const char* n) const;
    const Node* next_sibling(const char* n) const;
    const char* attribute(const char* n) const;  // "" when absent
};

class Document {
public:
    bool load_file(const char* path);
    bool parse(const std::string& src);
    const Node* child(const char* n) const { return root_.child(n); }
    const std::string& error() const { return error_; }

private:
    Node root_;
    std::string error_;
};

}  // namespace xml_lite
