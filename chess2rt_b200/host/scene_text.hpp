// Scene-description text front end: .sdl (SDLang subset) and .json -> abstract description nodes.
//
// Mirrors the *interface* the reference's loader is written against
// (/root/reference/source/rt/scene_loader.d:210-241 `SceneDscNode`, :243-331 `JsonValueWrapper`,
// :333-403 `SdlValueWrapper`): getType / getName / isSpecified / getChild / getChildren /
// getValues / get<T>.  The text parsers themselves (sdlang-d 0.10.6 and std.json in the
// reference) are replaced by two small hand-written recursive-descent parsers that cover the
// surface the bundled scenes use: tags with values and `{}` children,
// numbers, strings, booleans, `//`, `#`, `--` and `/* */` comments, `;` separators.
//
// This is host-side load-time plumbing, not the render path.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace c2rt_text {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Value {
    enum Kind { Null, Bool, Int, Float, String } kind = Null;
    bool b = false;
    long long i = 0;
    double f = 0;
    std::string s;

    bool isNumber() const { return kind == Int || kind == Float; }
    // sdlang's Variant.get!double accepts an int payload (implicit conversion);
    // get!long on a floating payload throws (scene_loader.d:390-398).
    double asDouble() const {
        if (kind == Int) return (double)i;
        if (kind == Float) return f;
        throw ParseError("value is not a number");
    }
    long long asInt() const {
        if (kind == Int) return i;
        throw ParseError("value is not an integer");
    }
    bool asBool() const {
        if (kind == Bool) return b;
        throw ParseError("value is not a boolean");
    }
    const std::string& asString() const {
        if (kind == String) return s;
        throw ParseError("value is not a string");
    }
};

// ---------------------------------------------------------------- SDLang subset
struct Tag {
    std::string name;
    std::vector<Value> values;
    std::vector<Tag> tags;

    const Tag* find(const std::string& n) const {
        for (auto& t : tags)
            if (t.name == n) return &t;
        return nullptr;
    }
};

class SdlParser {
public:
    explicit SdlParser(const std::string& src) : s_(src) {}

    Tag parseRoot() {
        Tag root;
        root.name = "root";
        parseTags(root, /*toplevel=*/true);
        return root;
    }

private:
    const std::string& s_;
    size_t p_ = 0;
    int line_ = 1;

    [[noreturn]] void fail(const std::string& m) const {
        throw ParseError("SDL line " + std::to_string(line_) + ": " + m);
    }
    bool eof() const { return p_ >= s_.size(); }
    char cur() const { return eof() ? '\0' : s_[p_]; }
    char peek(size_t k = 1) const { return p_ + k < s_.size() ? s_[p_ + k] : '\0'; }
    void adv() {
        if (cur() == '\n') line_++;
        p_++;
    }

    // skips blanks and comments; returns true if a tag terminator (newline / ';') was crossed
    bool skipSpace(bool crossNewlines) {
        bool crossed = false;
        for (;;) {
            char c = cur();
            if (c == ' ' || c == '\t' || c == '\r') { adv(); continue; }
            if (c == '\\' && (peek() == '\n' || (peek() == '\r' && peek(2) == '\n'))) {  // line continuation
                while (cur() != '\n') adv();
                adv();
                continue;
            }
            if (c == '/' && peek() == '*') {
                adv(); adv();
                while (!eof() && !(cur() == '*' && peek() == '/')) adv();
                if (eof()) fail("unterminated block comment");
                adv(); adv();
                continue;
            }
            if ((c == '/' && peek() == '/') || c == '#' || (c == '-' && peek() == '-')) {
                while (!eof() && cur() != '\n') adv();
                continue;
            }
            if (c == '\n' || c == ';') {
                crossed = true;
                if (!crossNewlines) return true;
                adv();
                continue;
            }
            return crossed;
        }
    }

    static bool identStart(char c) { return std::isalpha((unsigned char)c) || c == '_'; }
    static bool identChar(char c) {
        return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == '.' || c == '$' || c == ':';
    }

    std::string parseIdent() {
        size_t b = p_;
        while (identChar(cur())) adv();
        return s_.substr(b, p_ - b);
    }

    Value parseString() {
        Value v;
        v.kind = Value::String;
        char q = cur();
        adv();
        while (!eof() && cur() != q) {
            if (q == '"' && cur() == '\\') {
                adv();
                char e = cur();
                switch (e) {
                    case 'n': v.s += '\n'; break;
                    case 't': v.s += '\t'; break;
                    case 'r': v.s += '\r'; break;
                    case '0': v.s += '\0'; break;
                    default: v.s += e; break;
                }
                adv();
                continue;
            }
            v.s += cur();
            adv();
        }
        if (eof()) fail("unterminated string");
        adv();
        return v;
    }

    Value parseNumber() {
        size_t b = p_;
        if (cur() == '-' || cur() == '+') adv();
        bool isFloat = false;
        while (std::isdigit((unsigned char)cur())) adv();
        if (cur() == '.' && std::isdigit((unsigned char)peek())) {
            isFloat = true;
            adv();
            while (std::isdigit((unsigned char)cur())) adv();
        }
        if ((cur() == 'e' || cur() == 'E') &&
            (std::isdigit((unsigned char)peek()) || ((peek() == '-' || peek() == '+') && std::isdigit((unsigned char)peek(2))))) {
            isFloat = true;
            adv();
            if (cur() == '-' || cur() == '+') adv();
            while (std::isdigit((unsigned char)cur())) adv();
        }
        std::string tok = s_.substr(b, p_ - b);
        if (tok.empty() || tok == "-" || tok == "+") fail("malformed number");
        // type suffixes: L (long), f/F (float), d/D (double), BD (decimal)
        if (cur() == 'L') adv();
        else if (cur() == 'f' || cur() == 'F' || cur() == 'd' || cur() == 'D') { isFloat = true; adv(); }
        else if (cur() == 'B' && peek() == 'D') { isFloat = true; adv(); adv(); }
        Value v;
        if (isFloat) { v.kind = Value::Float; v.f = std::strtod(tok.c_str(), nullptr); }
        else { v.kind = Value::Int; v.i = std::strtoll(tok.c_str(), nullptr, 10); }
        return v;
    }

    bool tryParseValue(Value& out) {
        char c = cur();
        if (c == '"' || c == '`') { out = parseString(); return true; }
        if (std::isdigit((unsigned char)c) || ((c == '-' || c == '+') && (std::isdigit((unsigned char)peek()) || peek() == '.')) ||
            (c == '.' && std::isdigit((unsigned char)peek()))) {
            out = parseNumber();
            return true;
        }
        if (identStart(c)) {
            size_t save = p_;
            int saveLine = line_;
            std::string id = parseIdent();
            if (cur() != '=') {
                if (id == "true" || id == "on") { out.kind = Value::Bool; out.b = true; return true; }
                if (id == "false" || id == "off") { out.kind = Value::Bool; out.b = false; return true; }
                if (id == "null") { out.kind = Value::Null; return true; }
            }
            p_ = save;
            line_ = saveLine;
        }
        return false;
    }

    void parseTags(Tag& parent, bool toplevel) {
        for (;;) {
            skipSpace(true);
            if (eof()) {
                if (!toplevel) fail("missing '}'");
                return;
            }
            if (cur() == '}') {
                if (toplevel) fail("unexpected '}'");
                adv();
                return;
            }
            Tag t;
            if (identStart(cur())) {
                // could be a bare boolean/null value of an anonymous tag; the bundled scenes never do that
                t.name = parseIdent();
            } else {
                t.name = "content";  // sdlang's name for anonymous tags
            }
            // values, then attributes (attributes are parsed and dropped: the loader never reads them)
            for (;;) {
                bool ended = skipSpace(false);
                if (ended || eof() || cur() == '{' || cur() == '}') break;
                Value v;
                if (tryParseValue(v)) { t.values.push_back(v); continue; }
                if (identStart(cur())) {
                    parseIdent();
                    if (cur() != '=') fail("expected '=' after attribute name");
                    adv();
                    Value dummy;
                    if (!tryParseValue(dummy)) fail("malformed attribute value");
                    continue;
                }
                fail(std::string("unexpected character '") + cur() + "'");
            }
            if (cur() == '{') {
                adv();
                parseTags(t, false);
            }
            parent.tags.push_back(std::move(t));
        }
    }
};

// ---------------------------------------------------------------- JSON
struct Json {
    enum Kind { Null, Bool, Int, Float, String, Array, Object } kind = Null;
    bool b = false;
    long long i = 0;
    double f = 0;
    std::string s;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    const Json* find(const std::string& k) const {
        for (auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

class JsonParser {
public:
    explicit JsonParser(const std::string& src) : s_(src) {}
    Json parse() {
        Json j = parseValue();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return j;
    }

private:
    const std::string& s_;
    size_t p_ = 0;
    [[noreturn]] void fail(const std::string& m) const {
        throw ParseError("JSON offset " + std::to_string(p_) + ": " + m);
    }
    void ws() {
        while (p_ < s_.size() && std::isspace((unsigned char)s_[p_])) p_++;
    }
    char cur() const { return p_ < s_.size() ? s_[p_] : '\0'; }
    Json parseValue() {
        ws();
        Json j;
        char c = cur();
        if (c == '{') {
            j.kind = Json::Object;
            p_++;
            ws();
            if (cur() == '}') { p_++; return j; }
            for (;;) {
                ws();
                if (cur() != '"') fail("expected member name");
                std::string k = parseStr();
                ws();
                if (cur() != ':') fail("expected ':'");
                p_++;
                j.obj.emplace_back(k, parseValue());
                ws();
                if (cur() == ',') { p_++; continue; }
                if (cur() == '}') { p_++; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            j.kind = Json::Array;
            p_++;
            ws();
            if (cur() == ']') { p_++; return j; }
            for (;;) {
                j.arr.push_back(parseValue());
                ws();
                if (cur() == ',') { p_++; continue; }
                if (cur() == ']') { p_++; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') {
            j.kind = Json::String;
            j.s = parseStr();
        } else if (s_.compare(p_, 4, "true") == 0) { j.kind = Json::Bool; j.b = true; p_ += 4; }
        else if (s_.compare(p_, 5, "false") == 0) { j.kind = Json::Bool; j.b = false; p_ += 5; }
        else if (s_.compare(p_, 4, "null") == 0) { j.kind = Json::Null; p_ += 4; }
        else {
            size_t b = p_;
            bool isFloat = false;
            if (cur() == '-') p_++;
            while (std::isdigit((unsigned char)cur())) p_++;
            if (cur() == '.') { isFloat = true; p_++; while (std::isdigit((unsigned char)cur())) p_++; }
            if (cur() == 'e' || cur() == 'E') {
                isFloat = true; p_++;
                if (cur() == '-' || cur() == '+') p_++;
                while (std::isdigit((unsigned char)cur())) p_++;
            }
            if (p_ == b) fail("unexpected character");
            std::string tok = s_.substr(b, p_ - b);
            if (isFloat) { j.kind = Json::Float; j.f = std::strtod(tok.c_str(), nullptr); }
            else { j.kind = Json::Int; j.i = std::strtoll(tok.c_str(), nullptr, 10); }
        }
        return j;
    }
    std::string parseStr() {
        std::string out;
        p_++;
        while (p_ < s_.size() && s_[p_] != '"') {
            if (s_[p_] == '\\' && p_ + 1 < s_.size()) {
                p_++;
                switch (s_[p_]) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    default: out += s_[p_]; break;
                }
                p_++;
                continue;
            }
            out += s_[p_++];
        }
        if (p_ >= s_.size()) fail("unterminated string");
        p_++;
        return out;
    }
};

// ---------------------------------------------------------------- abstract description node
class DscNode {
public:
    virtual ~DscNode() = default;
    virtual std::string getType() const = 0;
    virtual bool hasName() const = 0;
    virtual std::string getName() const = 0;
    virtual bool isSpecified(const std::string& prop) const = 0;
    virtual std::unique_ptr<DscNode> getChild(const std::string& prop) const = 0;
    virtual std::vector<std::unique_ptr<DscNode>> getChildren() const = 0;
    virtual std::vector<Value> getValues() const = 0;
    virtual bool getBool() const = 0;
    virtual long long getInt() const = 0;
    virtual double getFloat() const = 0;
    virtual std::string getString() const = 0;
};

class SdlNode final : public DscNode {
public:
    explicit SdlNode(const Tag* t) : tag_(t) {}
    std::string getType() const override { return tag_->name; }
    bool hasName() const override {
        return (!tag_->values.empty() && tag_->values[0].kind == Value::String) || isSpecified("name");
    }
    std::string getName() const override {
        if (!tag_->values.empty() && tag_->values[0].kind == Value::String) return tag_->values[0].s;
        if (isSpecified("name")) return getChild("name")->getString();
        return std::string();
    }
    bool isSpecified(const std::string& prop) const override { return tag_->find(prop) != nullptr; }
    std::unique_ptr<DscNode> getChild(const std::string& prop) const override {
        const Tag* t = tag_->find(prop);
        if (!t) throw ParseError("missing property '" + prop + "'");
        return std::unique_ptr<DscNode>(new SdlNode(t));
    }
    std::vector<std::unique_ptr<DscNode>> getChildren() const override {
        std::vector<std::unique_ptr<DscNode>> r;
        for (auto& t : tag_->tags) r.emplace_back(new SdlNode(&t));
        return r;
    }
    std::vector<Value> getValues() const override { return tag_->values; }
    bool getBool() const override { return first().asBool(); }
    long long getInt() const override { return first().asInt(); }
    double getFloat() const override { return first().asDouble(); }
    std::string getString() const override { return first().asString(); }

private:
    const Tag* tag_;
    const Value& first() const {
        if (tag_->values.empty()) throw ParseError("tag '" + tag_->name + "' has no value");
        return tag_->values[0];
    }
};

class JsonNode final : public DscNode {
public:
    explicit JsonNode(const Json* j) : j_(j) {}
    std::string getType() const override {
        const Json* t = j_->find("type");
        if (!t || t->kind != Json::String) throw ParseError("JSON object has no \"type\"");
        return t->s;
    }
    bool hasName() const override { return isSpecified("name"); }
    std::string getName() const override {
        const Json* n = j_->find("name");
        return n ? n->s : std::string();
    }
    bool isSpecified(const std::string& prop) const override {
        return j_->kind == Json::Object && j_->find(prop) != nullptr;
    }
    std::unique_ptr<DscNode> getChild(const std::string& prop) const override {
        const Json* c = j_->find(prop);
        if (!c) throw ParseError("missing property '" + prop + "'");
        return std::unique_ptr<DscNode>(new JsonNode(c));
    }
    std::vector<std::unique_ptr<DscNode>> getChildren() const override {
        std::vector<std::unique_ptr<DscNode>> r;
        for (auto& c : j_->arr) r.emplace_back(new JsonNode(&c));
        return r;
    }
    std::vector<Value> getValues() const override {
        std::vector<Value> r;
        for (auto& c : j_->arr) {
            Value v;
            v.kind = Value::Float;
            v.f = number(c);
            r.push_back(v);
        }
        return r;
    }
    bool getBool() const override {
        if (j_->kind != Json::Bool) throw ParseError("JSON value is not a boolean");
        return j_->b;
    }
    long long getInt() const override { return (long long)number(*j_); }
    double getFloat() const override { return number(*j_); }
    std::string getString() const override {
        if (j_->kind != Json::String) throw ParseError("JSON value is not a string");
        return j_->s;
    }

private:
    const Json* j_;
    static double number(const Json& j) {
        if (j.kind == Json::Float) return j.f;
        if (j.kind == Json::Int) return (double)j.i;
        throw ParseError("JSON value is not a number");
    }
};

// Owns the parsed document and hands out the root description node
// (scene_loader.d:47-60: `.sdl` -> tags[0], `.json` -> the top-level object).
class Document {
public:
    static std::string readFile(const std::string& path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw ParseError("cannot open '" + path + "'");
        std::stringstream ss;
        ss << f.rdbuf();
        return ss.str();
    }
    static std::string lowerExt(const std::string& path) {
        size_t dot = path.find_last_of('.');
        size_t slash = path.find_last_of('/');
        if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
        std::string e = path.substr(dot);
        for (auto& c : e) c = (char)std::tolower((unsigned char)c);
        return e;
    }
    static std::string dirName(const std::string& path) {
        size_t slash = path.find_last_of('/');
        if (slash == std::string::npos) return ".";
        if (slash == 0) return "/";
        return path.substr(0, slash);
    }

    explicit Document(const std::string& path) {
        std::string ext = lowerExt(path);
        std::string text = readFile(path);
        if (ext == ".json") {
            json_ = JsonParser(text).parse();
            if (json_.kind != Json::Object) throw ParseError("top-level JSON value is not an object");
            isJson_ = true;
        } else if (ext == ".sdl") {
            sdl_ = SdlParser(text).parseRoot();
            if (sdl_.tags.empty()) throw ParseError("SDL file has no top-level tag");
        } else {
            throw ParseError("Error loading scene: unknown file type!");
        }
    }
    std::unique_ptr<DscNode> root() const {
        if (isJson_) return std::unique_ptr<DscNode>(new JsonNode(&json_));
        return std::unique_ptr<DscNode>(new SdlNode(&sdl_.tags[0]));
    }

private:
    bool isJson_ = false;
    Json json_;
    Tag sdl_;
};

}  // namespace c2rt_text
