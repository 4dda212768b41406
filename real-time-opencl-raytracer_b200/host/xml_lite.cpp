#include "xml_lite.h"

#include <cstdio>
#include <cstring>

namespace xml_lite {

const Node* Node::child(const char* n) const {
    for (const auto& c : children)
        if (c->name == n) return c.get();
    return nullptr;
}

const Node* Node::next_sibling(const char* n) const {
    if (!parent) return nullptr;
    const auto& sibs = parent->children;
    for (size_t i = index_in_parent + 1; i < sibs.size(); i++)
        if (sibs[i]->name == n) return sibs[i].get();
    return nullptr;
}

const char* Node::attribute(const char* n) const {
    for (const auto& a : attrs)
        if (a.first == n) return a.second.c_str();
    return "";
}

static void decode_entities(std::string& s) {
    if (s.find('&') == std::string::npos) return;
    static const struct { const char* from; char to; } ents[] = {
        {"&lt;", '<'}, {"&gt;", '>'}, {"&amp;", '&'}, {"&quot;", '"'}, {"&apos;", '\''}};
    std::string out;
    out.reserve(s.size());
    for (size_t i = 0; i < s.size();) {
        bool hit = false;
        if (s[i] == '&')
            for (const auto& e : ents) {
                size_t n = strlen(e.from);
                if (s.compare(i, n, e.from) == 0) {
                    out.push_back(e.to);
                    i += n;
                    hit = true;
                    break;
                }
            }
        if (!hit) out.push_back(s[i++]);
    }
    s.swap(out);
}

bool Document::load_file(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) {
        error_ = std::string("cannot open ") + path;
        return false;
    }
    std::string src;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) src.append(buf, n);
    fclose(f);
    return parse(src);
}

bool Document::parse(const std::string& s) {
    root_ = Node();
    error_.clear();
    Node* cur = &root_;
    size_t i = 0;
    const size_t n = s.size();
    auto is_space = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; };
    auto is_name = [&](char c) { return !is_space(c) && c != '>' && c != '/' && c != '=' && c != '<'; };
    while (i < n) {
        if (s[i] != '<') {  // character data
            size_t j = s.find('<', i);
            if (j == std::string::npos) j = n;
            if (cur != &root_ && !cur->text_set) {
                bool blank = true;
                for (size_t k = i; k < j && blank; k++) blank = is_space(s[k]);
                if (!blank) {
                    cur->text.assign(s, i, j - i);
                    decode_entities(cur->text);
                    cur->text_set = true;
                }
            }
            i = j;
            continue;
        }
        if (s.compare(i, 4, "<!--") == 0) {
            size_t j = s.find("-->", i + 4);
            if (j == std::string::npos) { error_ = "unterminated comment"; return false; }
            i = j + 3;
            continue;
        }
        if (s.compare(i, 9, "<![CDATA[") == 0) {
            size_t j = s.find("]]>", i + 9);
            if (j == std::string::npos) { error_ = "unterminated CDATA"; return false; }
            if (cur != &root_ && !cur->text_set) {
                cur->text.assign(s, i + 9, j - (i + 9));
                cur->text_set = true;
            }
            i = j + 3;
            continue;
        }
        if (i + 1 < n && (s[i + 1] == '?' || s[i + 1] == '!')) {  // declaration / doctype
            size_t j = s.find('>', i);
            if (j == std::string::npos) { error_ = "unterminated declaration"; return false; }
            i = j + 1;
            continue;
        }
        if (i + 1 < n && s[i + 1] == '/') {  // closing tag
            size_t j = s.find('>', i);
            if (j == std::string::npos) { error_ = "unterminated closing tag"; return false; }
            size_t a = i + 2, b = j;
            while (b > a && is_space(s[b - 1])) b--;
            if (cur == &root_ || s.compare(a, b - a, cur->name) != 0) {
                error_ = "mismatched closing tag </" + s.substr(a, b - a) + ">";
                return false;
            }
            cur = cur->parent;
            i = j + 1;
            continue;
        }
        // opening tag
        size_t j = i + 1;
        while (j < n && is_name(s[j])) j++;
        std::unique_ptr<Node> node(new Node());
        node->name.assign(s, i + 1, j - (i + 1));
        node->parent = cur;
        bool self_closing = false;
        for (;;) {
            while (j < n && is_space(s[j])) j++;
            if (j >= n) { error_ = "unterminated tag <" + node->name; return false; }
            if (s[j] == '>') { j++; break; }
            if (s[j] == '/') {
                self_closing = true;
                j++;
                continue;
            }
            size_t a = j;
            while (j < n && is_name(s[j])) j++;
            std::string key(s, a, j - a);
            while (j < n && is_space(s[j])) j++;
            std::string val;
            if (j < n && s[j] == '=') {
                j++;
                while (j < n && is_space(s[j])) j++;
                if (j < n && (s[j] == '"' || s[j] == '\'')) {
                    char q = s[j++];
                    size_t e = s.find(q, j);
                    if (e == std::string::npos) { error_ = "unterminated attribute value"; return false; }
                    val.assign(s, j, e - j);
                    decode_entities(val);
                    j = e + 1;
                }
            }
            if (key.empty()) { error_ = "malformed tag <" + node->name; return false; }
            node->attrs.emplace_back(std::move(key), std::move(val));
        }
        Node* raw = node.get();
        raw->index_in_parent = cur->children.size();
        cur->children.push_back(std::move(node));
        if (!self_closing) cur = raw;
        i = j;
    }
    if (cur != &root_) {
        error_ = "unclosed element <" + cur->name + ">";
        return false;
    }
    return true;
}

}  // namespace xml_lite
