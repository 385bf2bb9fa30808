#include "../include/ConfigParser.h"

#include <cctype>
#include <fstream>
#include <iostream>

namespace {

using Node = ConfigParser::Node;

struct Lexer {
    const std::string &s;
    size_t p = 0;
    int line = 1;
    explicit Lexer(const std::string &text) : s(text) {}
    // skips blanks and ';' comments; stops at newline if stopAtEol
    void skip(bool stopAtEol) {
        while (p < s.size()) {
            char c = s[p];
            if (c == '\n') {
                if (stopAtEol) return;
                ++line;
                ++p;
            } else if (c == ' ' || c == '\t' || c == '\r') {
                ++p;
            } else if (c == ';') {
                while (p < s.size() && s[p] != '\n') ++p;
            } else {
                return;
            }
        }
    }
    std::string token() { // bare word or "quoted string" (with \" \\ \n escapes; adjacent strings on
        std::string out;  // continuation lines `\` are not used by the demonstrator's files)
        if (s[p] == '"') {
            ++p;
            while (p < s.size() && s[p] != '"') {
                if (s[p] == '\\' && p + 1 < s.size()) {
                    ++p;
                    out += s[p] == 'n' ? '\n' : (s[p] == 't' ? '\t' : s[p]);
                } else {
                    if (s[p] == '\n') ++line;
                    out += s[p];
                }
                ++p;
            }
            if (p >= s.size()) throw std::runtime_error("config: unterminated string in line " + std::to_string(line));
            ++p;
        } else {
            while (p < s.size() && !std::isspace((unsigned char)s[p]) && s[p] != ';' && s[p] != '{' && s[p] != '}') out += s[p++];
        }
        return out;
    }
};

void parseInfoBlock(Lexer &lx, Node &node, bool top) {
    for (;;) {
        lx.skip(false);
        if (lx.p >= lx.s.size()) {
            if (!top) throw std::runtime_error("config: unmatched '{'");
            return;
        }
        if (lx.s[lx.p] == '}') {
            if (top) throw std::runtime_error("config: unmatched '}' in line " + std::to_string(lx.line));
            ++lx.p;
            return;
        }
        std::string key = lx.token();
        auto child = std::make_shared<Node>();
        lx.skip(true);
        if (lx.p < lx.s.size() && lx.s[lx.p] != '\n' && lx.s[lx.p] != '{' && lx.s[lx.p] != '}') child->value = lx.token();
        lx.skip(true);
        if (lx.p < lx.s.size() && lx.s[lx.p] == '\n') lx.skip(false); // the '{' may follow on the next line
        if (lx.p < lx.s.size() && lx.s[lx.p] == '{') {
            ++lx.p;
            parseInfoBlock(lx, *child, false);
        }
        node.children.emplace_back(key, child);
    }
}

// flat JSON: objects, arrays, strings, numbers/literals (enough for a config.json with the same keys)
struct Json {
    const std::string &s;
    size_t p = 0;
    explicit Json(const std::string &t) : s(t) {}
    void ws() {
        while (p < s.size() && std::isspace((unsigned char)s[p])) ++p;
    }
    std::string str() {
        std::string out;
        ++p;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) ++p;
            out += s[p++];
        }
        ++p;
        return out;
    }
    void value(Node &n) {
        ws();
        if (p >= s.size()) throw std::runtime_error("config: unexpected end of JSON");
        if (s[p] == '{') {
            ++p;
            for (;;) {
                ws();
                if (s[p] == '}') { ++p; return; }
                if (s[p] == ',') { ++p; continue; }
                if (s[p] != '"') throw std::runtime_error("config: JSON key expected");
                std::string k = str();
                ws();
                if (s[p] != ':') throw std::runtime_error("config: ':' expected in JSON");
                ++p;
                auto c = std::make_shared<Node>();
                value(*c);
                n.children.emplace_back(k, c);
            }
        } else if (s[p] == '[') {
            ++p;
            for (;;) {
                ws();
                if (s[p] == ']') { ++p; return; }
                if (s[p] == ',') { ++p; continue; }
                auto c = std::make_shared<Node>();
                value(*c);
                n.children.emplace_back("", c);
            }
        } else if (s[p] == '"') {
            n.value = str();
        } else {
            while (p < s.size() && s[p] != ',' && s[p] != '}' && s[p] != ']' && !std::isspace((unsigned char)s[p])) n.value += s[p++];
        }
    }
};

} // namespace

ConfigParser::ConfigParser(const std::string &file) : tree(std::make_shared<Node>()) {
    const size_t dot = file.find_last_of('.');
    const std::string ext = dot == std::string::npos ? "" : file.substr(dot);
    if (ext != ".info" && ext != ".json") {
        std::cerr << "Unsupported file extension: " << ext << std::endl; // as the reference: empty tree, lookups throw
        return;
    }
    std::ifstream in(file);
    if (!in) throw std::runtime_error(file + ": cannot open file");
    std::stringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    if (ext == ".json") {
        Json j(text);
        j.value(*tree);
    } else {
        Lexer lx(text);
        parseInfoBlock(lx, *tree, true);
    }
}

const ConfigParser::Node &ConfigParser::find(const std::string &path) const {
    const Node *n = tree.get();
    size_t at = 0;
    while (at <= path.size() && !path.empty()) {
        size_t dot = path.find('.', at);
        std::string part = path.substr(at, dot == std::string::npos ? std::string::npos : dot - at);
        const Node *next = nullptr;
        for (const auto &kv : n->children)
            if (kv.first == part) {
                next = kv.second.get();
                break;
            }
        if (!next) throw std::runtime_error("No such node (" + path + ")");
        n = next;
        if (dot == std::string::npos) break;
        at = dot + 1;
    }
    return *n;
}

ConfigParser ConfigParser::getObj(const std::string &key) {
    const Node &n = find(key);
    auto copy = std::make_shared<Node>(n);
    return ConfigParser(copy);
}

std::list<ConfigParser> ConfigParser::getObjList(const std::string &key) {
    std::list<ConfigParser> out;
    for (const auto &kv : find(key).children) {
        if (kv.second->children.empty()) {
            std::cerr << "List does not contain objects. Please use 'getList<T>(const std::string &key)'instead."
                      << " - Returning empty list." << std::endl;
        } else {
            out.push_back(ConfigParser(kv.second));
        }
    }
    return out;
}
