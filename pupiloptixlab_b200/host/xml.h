// Pupil::resource::xml — the mitsuba3-style scene dialect of the reference, parsed by a small hand-written
// XML reader (the reference uses pugixml).
//
//   Object / Property / GlobalManager   framework/resource/xml/object.{h,cpp}
//   tag set and per-tag visitors        framework/resource/xml/tag.h:11-33, visitor.h:18-207
//   Parser::LoadFromFile (DFS)          framework/resource/xml/parser.cpp:22-58
//
// Dialect rules kept: <default name value> + "$name" substitution in every attribute value (applied when a
// node is visited, so a default only affects nodes after it); objects (bsdf emitter film integrator sensor
// shape texture transform) become Object nodes, `id` registers them for <ref id>; property tags (integer
// string float rgb boolean matrix) become (name, value) pairs where a missing `name` falls back to the tag
// name; point / scale / translate accept value="a,b,c" or x= y= z= with per-tag defaults; <rotate> keeps
// axis + angle as a sub object; <lookat> keeps origin / target / up; unknown tags are skipped with their
// whole subtree (visitor.h:82-87).
#pragma once
#include <memory>
#include <string>
#include <string_view>
#include <unordered_map>
#include <utility>
#include <vector>

namespace Pupil::resource::xml {

enum class ETag : unsigned {
    _unknown = 0, _scene, _default, _bsdf, _emitter, _film, _integrator, _sensor, _shape, _texture, _lookat, _transform,
    _integer, _string, _float, _rgb, _point, _matrix, _scale, _rotate, _translate, _boolean, _ref, _count
};

struct Property {
    std::string name, value;
};

struct Object {
    ETag tag = ETag::_unknown;
    std::string obj_name; // the tag name
    std::string var_name; // name="..."
    std::string id, type;
    std::vector<Property> properties;
    std::vector<Object *> sub_object;

    std::string GetProperty(std::string_view name) const noexcept;
    Object *GetUniqueSubObject(std::string_view obj_name) const noexcept;
    std::vector<Object *> GetSubObjects(std::string_view obj_name) const noexcept;
    // a parameter is either a named sub object (a <texture name="...">) or a plain property
    std::pair<Object *, std::string> GetParameter(std::string_view name) const noexcept;
};

// raw XML element as read from the file
struct Node {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<Node> children;
    const std::string *Attr(std::string_view key) const noexcept {
        for (auto &a : attrs)
            if (a.first == key) return &a.second;
        return nullptr;
    }
    std::string AttrOr(std::string_view key, std::string_view fallback = "") const {
        auto *v = Attr(key);
        return v ? *v : std::string(fallback);
    }
};

// Parses a document; returns false (and a message) on malformed input.  Handles the prolog, comments,
// CDATA-free element trees, single/double quoted attributes and the five predefined entities.
bool ParseDocument(std::string_view text, Node &root, std::string &error);

class Parser {
public:
    // returns the <scene> object or nullptr
    Object *LoadFromFile(const std::string &path) noexcept;
    Object *LoadFromString(std::string_view text) noexcept;
    const std::string &Error() const noexcept { return m_error; }

private:
    void Visit(const Node &node);
    Object *NewObject(std::string_view name, std::string_view type, ETag tag);
    std::string Substitute(std::string value) const;

    std::vector<std::unique_ptr<Object>> m_pool;
    std::vector<std::pair<std::string, std::string>> m_defaults; // in declaration order
    std::unordered_map<std::string, Object *> m_refs;
    Object *m_current = nullptr;
    std::string m_error;
};
}// namespace Pupil::resource::xml
