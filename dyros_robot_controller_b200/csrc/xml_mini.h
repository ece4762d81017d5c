// drc_b200 -- minimal XML reader (elements, attributes, comments, declarations) for URDF / SRDF.
// Enough for robot description files; no entities beyond the five predefined ones, no DTD.
#pragma once
#include <cctype>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace drc {
namespace xml {

struct Node {
  std::string tag;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<std::unique_ptr<Node>> children;
  std::string text;  // character data directly inside the element (COLLADA float arrays; unused by URDF / SRDF)

  const std::string* attr(const std::string& k) const {
    for (auto& a : attrs) if (a.first == k) return &a.second;
    return nullptr;
  }
  std::string attr_or(const std::string& k, const std::string& d) const {
    const std::string* v = attr(k);
    return v ? *v : d;
  }
  const Node* child(const std::string& t) const {
    for (auto& c : children) if (c->tag == t) return c.get();
    return nullptr;
  }
  std::vector<const Node*> all(const std::string& t) const {
    std::vector<const Node*> r;
    for (auto& c : children) if (c->tag == t) r.push_back(c.get());
    return r;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s), i_(0) {}
  std::unique_ptr<Node> parse() {
    skip_misc();
    auto root = element();
    if (!root) throw std::runtime_error("xml: no root element");
    return root;
  }

 private:
  const std::string& s_;
  size_t i_;

  bool starts(const char* lit) const { return s_.compare(i_, std::char_traits<char>::length(lit), lit) == 0; }
  void skip_ws() { while (i_ < s_.size() && std::isspace((unsigned char)s_[i_])) ++i_; }
  void skip_until(const char* lit) {
    size_t p = s_.find(lit, i_);
    if (p == std::string::npos) throw std::runtime_error("xml: unterminated construct");
    i_ = p + std::char_traits<char>::length(lit);
  }
  void skip_misc() {
    for (;;) {
      skip_ws();
      if (starts("<?")) skip_until("?>");
      else if (starts("<!--")) skip_until("-->");
      else if (starts("<!")) skip_until(">");
      else break;
    }
  }
  static std::string unescape(const std::string& v) {
    std::string o;
    for (size_t k = 0; k < v.size(); ++k) {
      if (v[k] != '&') { o += v[k]; continue; }
      if (v.compare(k, 4, "&lt;") == 0) { o += '<'; k += 3; }
      else if (v.compare(k, 4, "&gt;") == 0) { o += '>'; k += 3; }
      else if (v.compare(k, 5, "&amp;") == 0) { o += '&'; k += 4; }
      else if (v.compare(k, 6, "&quot;") == 0) { o += '"'; k += 5; }
      else if (v.compare(k, 6, "&apos;") == 0) { o += '\''; k += 5; }
      else o += v[k];
    }
    return o;
  }
  std::string name() {
    size_t b = i_;
    while (i_ < s_.size() && (std::isalnum((unsigned char)s_[i_]) || s_[i_] == '_' || s_[i_] == '-' || s_[i_] == ':' || s_[i_] == '.')) ++i_;
    return s_.substr(b, i_ - b);
  }
  std::unique_ptr<Node> element() {
    if (i_ >= s_.size() || s_[i_] != '<') return nullptr;
    ++i_;
    std::unique_ptr<Node> n(new Node);
    n->tag = name();
    for (;;) {
      skip_ws();
      if (i_ >= s_.size()) throw std::runtime_error("xml: unexpected end in <" + n->tag + ">");
      if (s_[i_] == '/') {
        if (i_ + 1 >= s_.size() || s_[i_ + 1] != '>') throw std::runtime_error("xml: malformed tag <" + n->tag + ">");
        i_ += 2;
        return n;
      }
      if (s_[i_] == '>') { ++i_; break; }
      std::string k = name();
      if (k.empty()) throw std::runtime_error("xml: bad attribute in <" + n->tag + ">");
      skip_ws();
      if (i_ >= s_.size() || s_[i_] != '=') throw std::runtime_error("xml: attribute without value in <" + n->tag + ">");
      ++i_;
      skip_ws();
      if (i_ >= s_.size()) throw std::runtime_error("xml: unexpected end");
      char qc = s_[i_];
      if (qc != '"' && qc != '\'') throw std::runtime_error("xml: unquoted attribute in <" + n->tag + ">");
      size_t e = s_.find(qc, i_ + 1);
      if (e == std::string::npos) throw std::runtime_error("xml: unterminated attribute");
      n->attrs.emplace_back(k, unescape(s_.substr(i_ + 1, e - i_ - 1)));
      i_ = e + 1;
    }
    // children / text until the closing tag
    for (;;) {
      size_t lt = s_.find('<', i_);
      if (lt == std::string::npos) throw std::runtime_error("xml: missing </" + n->tag + ">");
      if (lt > i_) n->text.append(s_, i_, lt - i_);
      i_ = lt;
      if (starts("<!--")) { skip_until("-->"); continue; }
      if (starts("<![CDATA[")) { skip_until("]]>"); continue; }
      if (starts("<?")) { skip_until("?>"); continue; }
      if (starts("</")) {
        i_ += 2;
        std::string t = name();
        if (t != n->tag) throw std::runtime_error("xml: </" + t + "> closes <" + n->tag + ">");
        skip_ws();
        if (i_ >= s_.size() || s_[i_] != '>') throw std::runtime_error("xml: malformed closing tag");
        ++i_;
        return n;
      }
      n->children.push_back(element());
    }
  }
};

inline std::unique_ptr<Node> parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace xml
}  // namespace drc
