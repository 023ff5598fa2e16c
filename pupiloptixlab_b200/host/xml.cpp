#include "xml.h"
#include "util.h"

#include <algorithm>
#include <cctype>
#include <fstream>
#include <sstream>

namespace Pupil::resource::xml {

// ---- Object ---------------------------------------------------------------------------------------------
std::string Object::GetProperty(std::string_view name) const noexcept {
    for (auto &p : properties)
        if (p.name == name) return p.value;
    return "";
}
Object *Object::GetUniqueSubObject(std::string_view name) const noexcept {
    for (auto *so : sub_object)
        if (so->obj_name == name) return so;
    return nullptr;
}
std::vector<Object *> Object::GetSubObjects(std::string_view name) const noexcept {
    std::vector<Object *> out;
    for (auto *so : sub_object)
        if (so->obj_name == name) out.push_back(so);
    return out;
}
std::pair<Object *, std::string> Object::GetParameter(std::string_view name) const noexcept {
    for (auto *so : sub_object)
        if (so->var_name == name) return { so, "" };
    return { nullptr, GetProperty(name) };
}

// ---- a small recursive-descent XML reader -------------------------------------------------------------------
namespace {
struct Reader {
    std::string_view s;
    size_t i = 0;
    std::string err;

    bool Fail(const std::string &m) {
        if (err.empty()) err = m + " at byte " + std::to_string(i);
        return false;
    }
    bool StartsWith(std::string_view t) const { return s.substr(i, t.size()) == t; }
    void SkipSpace() {
        while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i;
    }
    bool SkipUntil(std::string_view end) {
        const size_t j = s.find(end, i);
        if (j == std::string_view::npos) return Fail("unterminated construct");
        i = j + end.size();
        return true;
    }
    // comments, processing instructions, doctype
    bool SkipMisc() {
        for (;;) {
            SkipSpace();
            if (StartsWith("<!--")) {
                if (!SkipUntil("-->")) return false;
            } else if (StartsWith("<?")) {
                if (!SkipUntil("?>")) return false;
            } else if (StartsWith("<!")) {
                if (!SkipUntil(">")) return false;
            } else
                return true;
        }
    }
    static bool NameChar(char c) { return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '-' || c == ':' || c == '.'; }
    std::string Name() {
        const size_t b = i;
        while (i < s.size() && NameChar(s[i])) ++i;
        return std::string(s.substr(b, i - b));
    }
    static std::string Unescape(std::string_view v) {
        std::string out;
        out.reserve(v.size());
        for (size_t k = 0; k < v.size(); ++k) {
            if (v[k] != '&') {
                out.push_back(v[k]);
                continue;
            }
            static const std::pair<std::string_view, char> ents[] = { { "&lt;", '<' }, { "&gt;", '>' }, { "&amp;", '&' }, { "&quot;", '"' }, { "&apos;", '\'' } };
            bool done = false;
            for (auto &[e, c] : ents)
                if (v.substr(k, e.size()) == e) {
                    out.push_back(c), k += e.size() - 1, done = true;
                    break;
                }
            if (!done && k + 2 < v.size() && v[k + 1] == '#') { // numeric character reference: &#NN; / &#xHH; -> UTF-8
                const bool hex = v[k + 2] == 'x' || v[k + 2] == 'X';
                size_t j = k + (hex ? 3 : 2);
                uint32_t cp = 0;
                size_t digits = 0;
                for (; j < v.size() && digits < 8; ++j, ++digits) {
                    const char c = v[j];
                    int d = (c >= '0' && c <= '9') ? c - '0' : (hex && c >= 'a' && c <= 'f') ? c - 'a' + 10 : (hex && c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1;
                    if (d < 0) break;
                    cp = cp * (hex ? 16u : 10u) + (uint32_t)d;
                }
                if (digits > 0 && j < v.size() && v[j] == ';' && cp > 0 && cp <= 0x10FFFFu) {
                    if (cp < 0x80) out.push_back((char)cp);
                    else if (cp < 0x800) out.push_back((char)(0xC0 | cp >> 6)), out.push_back((char)(0x80 | (cp & 0x3F)));
                    else if (cp < 0x10000) out.push_back((char)(0xE0 | cp >> 12)), out.push_back((char)(0x80 | (cp >> 6 & 0x3F))), out.push_back((char)(0x80 | (cp & 0x3F)));
                    else out.push_back((char)(0xF0 | cp >> 18)), out.push_back((char)(0x80 | (cp >> 12 & 0x3F))), out.push_back((char)(0x80 | (cp >> 6 & 0x3F))), out.push_back((char)(0x80 | (cp & 0x3F)));
                    k = j, done = true;
                }
            }
            if (!done) out.push_back('&');
        }
        return out;
    }
    // Element() recurses once per nesting level: the depth is capped so that a hostile file fails instead of overflowing
    // the stack (the dialect nests five or six levels deep)
    static constexpr int kMaxDepth = 256;
    int depth = 0;
    bool Element(Node &node) {
        if (depth >= kMaxDepth) return Fail("elements nested deeper than " + std::to_string(kMaxDepth) + " levels");
        ++depth;
        const bool ok = ElementBody(node);
        --depth;
        return ok;
    }
    bool ElementBody(Node &node) {
        if (i >= s.size() || s[i] != '<') return Fail("expected '<'");
        ++i;
        node.name = Name();
        if (node.name.empty()) return Fail("empty element name");
        for (;;) {
            SkipSpace();
            if (i >= s.size()) return Fail("unterminated start tag");
            if (s[i] == '/') {
                if (!StartsWith("/>")) return Fail("expected '/>'");
                i += 2;
                return true;
            }
            if (s[i] == '>') {
                ++i;
                break;
            }
            std::string key = Name();
            if (key.empty()) return Fail("bad attribute name");
            SkipSpace();
            if (i >= s.size() || s[i] != '=') return Fail("expected '='");
            ++i;
            SkipSpace();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) return Fail("expected a quoted attribute value");
            const char q = s[i++];
            const size_t j = s.find(q, i);
            if (j == std::string_view::npos) return Fail("unterminated attribute value");
            node.attrs.emplace_back(std::move(key), Unescape(s.substr(i, j - i)));
            i = j + 1;
        }
        // content
        for (;;) {
            const size_t lt = s.find('<', i);
            if (lt == std::string_view::npos) return Fail("missing end tag for <" + node.name + ">");
            i = lt; // character data is not part of the dialect
            if (StartsWith("<!--")) {
                if (!SkipUntil("-->")) return false;
            } else if (StartsWith("<![CDATA[")) {
                if (!SkipUntil("]]>")) return false;
            } else if (StartsWith("<?")) {
                if (!SkipUntil("?>")) return false;
            } else if (StartsWith("</")) {
                i += 2;
                const std::string end = Name();
                if (end != node.name) return Fail("mismatched end tag </" + end + "> for <" + node.name + ">");
                SkipSpace();
                if (i >= s.size() || s[i] != '>') return Fail("expected '>'");
                ++i;
                return true;
            } else {
                node.children.emplace_back();
                if (!Element(node.children.back())) return false;
            }
        }
    }
};

ETag TagOf(std::string_view name) {
    static const std::unordered_map<std::string_view, ETag> map = {
        { "scene", ETag::_scene }, { "default", ETag::_default }, { "bsdf", ETag::_bsdf }, { "emitter", ETag::_emitter }, { "film", ETag::_film },
        { "integrator", ETag::_integrator }, { "sensor", ETag::_sensor }, { "shape", ETag::_shape }, { "texture", ETag::_texture },
        { "lookat", ETag::_lookat }, { "transform", ETag::_transform }, { "integer", ETag::_integer }, { "string", ETag::_string },
        { "float", ETag::_float }, { "rgb", ETag::_rgb }, { "point", ETag::_point }, { "matrix", ETag::_matrix }, { "scale", ETag::_scale },
        { "rotate", ETag::_rotate }, { "translate", ETag::_translate }, { "boolean", ETag::_boolean }, { "ref", ETag::_ref }
    };
    auto it = map.find(name);
    return it == map.end() ? ETag::_unknown : it->second;
}
}// namespace

bool ParseDocument(std::string_view text, Node &root, std::string &error) {
    Reader r{ text };
    if (text.size() >= 3 && static_cast<unsigned char>(text[0]) == 0xEF && static_cast<unsigned char>(text[1]) == 0xBB) r.i = 3; // UTF-8 BOM
    if (!r.SkipMisc() || !r.Element(root)) {
        error = r.err.empty() ? "no root element" : r.err;
        return false;
    }
    return true;
}

// ---- scene-dialect visitors -----------------------------------------------------------------------------------
Object *Parser::NewObject(std::string_view name, std::string_view type, ETag tag) {
    auto obj = std::make_unique<Object>();
    obj->obj_name = name, obj->type = type, obj->tag = tag;
    m_pool.push_back(std::move(obj));
    return m_pool.back().get();
}

// "$name" -> value for every declared default (object.cpp:9-24); longer names first so "$resx" is not
// eaten by "$res"
std::string Parser::Substitute(std::string value) const {
    if (value.find('$') == std::string::npos) return value;
    std::vector<const std::pair<std::string, std::string> *> order;
    for (auto &d : m_defaults) order.push_back(&d);
    std::stable_sort(order.begin(), order.end(), [](auto *a, auto *b) { return a->first.size() > b->first.size(); });
    for (auto *d : order) {
        const std::string key = "$" + d->first;
        for (size_t pos = 0; (pos = value.find(key, pos)) != std::string::npos; pos += d->second.size()) value.replace(pos, key.size(), d->second);
    }
    return value;
}

void Parser::Visit(const Node &node) {
    const ETag tag = TagOf(node.name);
    Object *parent = m_current;
    auto attr = [&](std::string_view key) { return Substitute(node.AttrOr(key)); };
    auto has = [&](std::string_view key) { return node.Attr(key) != nullptr; };
    auto add_property = [&](std::string name, std::string value) {
        if (m_current) m_current->properties.push_back({ std::move(name), std::move(value) });
    };
    auto xyz_property = [&](const char *dx, const char *dy, const char *dz) { // visitor.h:57-76
        std::string name = attr("name");
        if (name.empty()) name = node.name;
        std::string value = attr("value");
        if (value.empty()) {
            std::string x = attr("x"), y = attr("y"), z = attr("z");
            value = (x.empty() ? dx : x) + "," + (y.empty() ? dy : y) + "," + (z.empty() ? dz : z);
        }
        add_property(std::move(name), std::move(value));
    };

    switch (tag) {
        case ETag::_scene: m_current = NewObject(node.name, attr("version"), tag); break;
        case ETag::_default: { // later declarations of the same name win
            const std::string name = attr("name"), value = attr("value");
            auto it = std::find_if(m_defaults.begin(), m_defaults.end(), [&](auto &d) { return d.first == name; });
            if (it != m_defaults.end()) it->second = value;
            else m_defaults.emplace_back(name, value);
        } break;
        case ETag::_ref:
            if (has("id") && m_current) {
                auto it = m_refs.find(attr("id"));
                if (it != m_refs.end()) m_current->sub_object.push_back(it->second);
            }
            break;
        case ETag::_lookat: {
            Object *o = NewObject(node.name, "", tag);
            o->properties.push_back({ "origin", attr("origin") });
            o->properties.push_back({ "target", attr("target") });
            o->properties.push_back({ "up", attr("up") });
            if (m_current) m_current->sub_object.push_back(o);
        } break;
        case ETag::_rotate: { // <rotate value="x,y,z" angle> | <rotate y="1" angle>
            Object *o = NewObject(node.name, "", tag);
            std::string axis;
            if (has("value")) axis = attr("value");
            else if (has("x")) axis = "1, 0, 0";
            else if (has("y")) axis = "0, 1, 0";
            else if (has("z")) axis = "0, 0, 1";
            o->properties.push_back({ "axis", axis });
            o->properties.push_back({ "angle", attr("angle") });
            if (m_current) m_current->sub_object.push_back(o);
        } break;
        case ETag::_scale: xyz_property("1", "1", "1"); break;
        case ETag::_point: xyz_property("0", "0", "0"); break;
        case ETag::_translate: xyz_property("0", "0", "0"); break;
        case ETag::_integer:
        case ETag::_string:
        case ETag::_float:
        case ETag::_rgb:
        case ETag::_boolean:
        case ETag::_matrix: {
            std::string name = attr("name");
            if (name.empty()) name = node.name;
            add_property(std::move(name), attr("value"));
        } break;
        case ETag::_bsdf:
        case ETag::_emitter:
        case ETag::_film:
        case ETag::_integrator:
        case ETag::_sensor:
        case ETag::_shape:
        case ETag::_texture:
        case ETag::_transform: {
            Object *o = NewObject(node.name, attr("type"), tag);
            if (has("id")) o->id = attr("id"), m_refs[o->id] = o;
            if (has("name")) o->var_name = attr("name");
            if (m_current) m_current->sub_object.push_back(o);
            m_current = o;
        } break;
        default:
            Log::Warn("XML node [%s] skipped", node.name.c_str());
            return; // the whole subtree
    }
    for (auto &ch : node.children) Visit(ch);
    m_current = parent;
}

Object *Parser::LoadFromString(std::string_view text) noexcept {
    m_pool.clear(), m_defaults.clear(), m_refs.clear(), m_current = nullptr, m_error.clear();
    Node root;
    if (!ParseDocument(text, root, m_error)) {
        Log::Error("XML parse error: %s", m_error.c_str());
        return nullptr;
    }
    Visit(root);
    if (m_pool.empty() || m_pool[0]->tag != ETag::_scene) {
        m_error = "the root element is not <scene>";
        return nullptr;
    }
    return m_pool[0].get();
}

Object *Parser::LoadFromFile(const std::string &path) noexcept {
    std::ifstream f(path, std::ios::binary);
    if (!f) {
        m_error = "cannot open " + path;
        return nullptr;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    return LoadFromString(text);
}
}// namespace Pupil::resource::xml
