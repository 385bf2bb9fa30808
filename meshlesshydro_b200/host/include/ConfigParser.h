// config.info reader with the interface of the reference's ConfigParser
// (/root/reference/demonstrator/include/ConfigParser.h:13-53), without Boost: a small parser for the
// property-tree INFO format the demonstrator's config files use (`key value`, `key { ... }` blocks,
// `;` comments, quoted strings) and for flat JSON objects.
#ifndef CONFIGPARSER_H
#define CONFIGPARSER_H

#include <list>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

class ConfigParser {
public:
    explicit ConfigParser(const std::string &file = "config.info");

    std::list<ConfigParser> getObjList(const std::string &key);
    ConfigParser getObj(const std::string &key);

    template <typename T> T getVal(const std::string &key);
    template <typename T> std::list<T> getList(const std::string &key);

    struct Node { // ordered tree: value + children (boost::property_tree::ptree stand-in)
        std::string value;
        std::vector<std::pair<std::string, std::shared_ptr<Node>>> children;
    };

private:
    explicit ConfigParser(std::shared_ptr<Node> subtree) : tree(std::move(subtree)) {}
    std::shared_ptr<Node> tree;
    const Node &find(const std::string &path) const; // dotted path; throws std::runtime_error("No such node (..)")
    template <typename T> static T convert(const std::string &text, const std::string &key);
};

template <typename T> T ConfigParser::convert(const std::string &text, const std::string &key) {
    std::istringstream is(text);
    T v;
    is >> v;
    if (is.fail() || (is >> std::ws, !is.eof()))
        throw std::runtime_error("conversion of data to type failed for key \"" + key + "\" (value \"" + text + "\")");
    return v;
}
template <> inline std::string ConfigParser::convert<std::string>(const std::string &text, const std::string &) { return text; }

template <typename T> T ConfigParser::getVal(const std::string &key) { return convert<T>(find(key).value, key); }

template <typename T> std::list<T> ConfigParser::getList(const std::string &key) {
    std::list<T> out;
    for (const auto &kv : find(key).children) {
        if (kv.second->children.empty()) {
            out.push_back(convert<T>(kv.second->value, key));
        } else {
            out.clear();
            break; // list of objects: use getObjList (the reference prints a hint and returns what it has)
        }
    }
    return out;
}

#endif // CONFIGPARSER_H
