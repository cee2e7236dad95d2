"""An INDEPENDENT reader of the SDLang subset the bundled scene files use — written for the tests only, sharing no code with
chess2rt_b200/host/scene_text.hpp (which both the host loader AND the oracle use: a parse bug there would be common-mode,
VERDICT r1).  Grammar covered: `Tag value* { children }`, `Tag value*` ended by newline or `;`, values = numbers, "strings",
true/false/on/off, comments `//`, `#`, `--`, `/* */`.  Returns nested (name, values, children) tuples."""
import re

TOKEN = re.compile(r'''\s*(?:(//[^\n]*|\#[^\n]*|--[^\n]*|/\*.*?\*/)|("(?:[^"\\]|\\.)*")|([{};])|(\n)|([^\s{};"]+))''', re.S)


def tokenize(text):
    pos, out = 0, []
    text = text.replace("\r\n", "\n")
    while pos < len(text):
        m = re.compile(r'[ \t]*').match(text, pos)
        pos = m.end()
        if pos >= len(text):
            break
        if text[pos] == "\n":
            out.append(("nl", "\n")); pos += 1; continue
        if text.startswith("//", pos) or text.startswith("#", pos) or text.startswith("--", pos):
            e = text.find("\n", pos)
            pos = len(text) if e < 0 else e
            continue
        if text.startswith("/*", pos):
            pos = text.index("*/", pos) + 2
            continue
        if text[pos] == '"':
            e = pos + 1
            while text[e] != '"':
                e += 2 if text[e] == "\\" else 1
            out.append(("str", text[pos + 1:e])); pos = e + 1; continue
        if text[pos] in "{};":
            out.append((text[pos], text[pos])); pos += 1; continue
        m = re.compile(r'[^\s{};"]+').match(text, pos)
        out.append(("word", m.group(0))); pos = m.end()
    return out


def value(kind, tok):
    if kind == "str":
        return tok
    low = tok.lower()
    if low in ("true", "on"):
        return True
    if low in ("false", "off"):
        return False
    try:
        return float(tok.rstrip("fFdD")) if re.search(r"[.eE]", tok) or tok[-1] in "fFdD" else int(tok.rstrip("lL"))
    except ValueError:
        return tok


def parse(text):
    toks = tokenize(text)
    i = 0

    def tags(depth):
        nonlocal i
        res = []
        while i < len(toks):
            kind, tok = toks[i]
            if kind in ("nl", ";"):
                i += 1
                continue
            if kind == "}":
                assert depth > 0
                i += 1
                return res
            assert kind == "word", (kind, tok)
            name, vals, kids = tok, [], []
            i += 1
            while i < len(toks) and toks[i][0] in ("word", "str"):
                vals.append(value(*toks[i]))
                i += 1
            if i < len(toks) and toks[i][0] == "{":
                i += 1
                kids = tags(depth + 1)
            res.append((name, vals, kids))
        return res

    return tags(0)


def child(tag, name):
    for t in tag[2]:
        if t[0] == name:
            return t
    return None


def prop(tag, name, default=None):
    c = child(tag, name)
    return default if c is None else (c[1][0] if len(c[1]) == 1 else c[1])


def obj_name(tag):   # `Sphere "s" {` or a `name "s"` child (scene_loader.d:333-403)
    if tag[1] and isinstance(tag[1][0], str):
        return tag[1][0]
    return prop(tag, "name")
