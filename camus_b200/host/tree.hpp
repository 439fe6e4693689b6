// Host-side tree model for the CAMUS drop-in: rooted n-ary trees, newick / nexus readers,
// and the gotree operations the reference calls before the hot path.
//
// Mirrors (behaviour, not code) the parts of github.com/evolbioinfo/gotree v0.4.5 that
// jsdoublel/camus uses, as restated in SURVEY.md §8(c) G1-G8:
//   newick parser           <- internal/prep/io.go:111,141   (newick.NewParser(..).Parse())
//   nexus TREES reader      <- internal/prep/io.go:157-164
//   UpdateTipIndex          <- internal/prep/preprocess.go:31,87      (G1: rank of label, byte order)
//   Nodes()/SetId preorder  <- internal/prep/preprocess.go:28-30      (G2: preorder, root = 0)
//   RemoveSingleNodes       <- internal/prep/preprocess.go:27
//   CollapseLowSupport      <- internal/prep/preprocess.go:98-100     (G6)
//   Rooted()/TreeIsBinary   <- internal/prep/preprocess.go:34-39,151-177
// This file is host plumbing (parsing/flattening); nothing here is on the GPU hot path.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <atomic>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

namespace camus {

// Sentinel error classes of the reference (preprocess.go:17-22, quartet.go:33-36, io.go:29-33).
enum class Err : int {
  Ok = 0,
  Unrooted = 1,         // prep.ErrUnrooted
  NonBinary = 2,        // prep.ErrNonBinary
  MulTree = 3,          // prep.ErrMulTree
  TipNameMismatch = 4,  // graphs.ErrTipNameMismatch
  InvalidFile = 5,      // prep.ErrInvalidFile
  InvalidFormat = 6,    // prep.ErrInvalidFormat
  TypeOutRange = 7,     // prep.ErrTypeOutRange
  InvalidScorerOption = 8,  // score.ErrInvalidScorerOption
  Device = 9,           // CUDA / C-ABI failure (no reference counterpart)
  Internal = 10,
};

struct CamusError : std::runtime_error {
  Err code;
  CamusError(Err c, const std::string& msg) : std::runtime_error(msg), code(c) {}
};

struct Node {
  std::string name;
  int parent = -1;
  std::vector<int> children;
  bool has_support = false;
  double support = 0.0;
  int tip_index = -1;  // gotree tip index (rank of the label); -1 for internal nodes
  bool tip() const { return children.empty(); }
};

// Rooted tree stored as an index pool. `order` lists live nodes in preorder; node "ids" in the
// sense of the reference (preprocess.go:28-30) are positions in that list.
struct Tree {
  std::vector<Node> nodes;
  int root = -1;

  int add_node() {
    nodes.emplace_back();
    return (int)nodes.size() - 1;
  }

  // Preorder walk from the root following child order (gotree Nodes()/PreOrder).
  std::vector<int> preorder() const {
    std::vector<int> out, stack;
    if (root < 0) return out;
    stack.push_back(root);
    while (!stack.empty()) {
      int x = stack.back();
      stack.pop_back();
      out.push_back(x);
      const auto& ch = nodes[x].children;
      for (auto it = ch.rbegin(); it != ch.rend(); ++it) stack.push_back(*it);
    }
    return out;
  }

  std::vector<int> postorder() const {
    std::vector<int> pre = preorder_mirrored();
    std::reverse(pre.begin(), pre.end());
    return pre;
  }

  std::vector<int> tips() const {
    std::vector<int> out;
    for (int x : preorder())
      if (nodes[x].tip()) out.push_back(x);
    return out;
  }

  int n_tips() const { return (int)tips().size(); }

  // gotree Rooted(): the root has exactly two neighbours.
  bool rooted() const { return root >= 0 && nodes[root].children.size() == 2; }

  // prep.TreeIsBinary (preprocess.go:151-160): rooted and every internal node has two children.
  bool is_binary() const {
    if (!rooted()) return false;
    for (int x : preorder())
      if (!nodes[x].tip() && nodes[x].children.size() != 2) return false;
    return true;
  }

  // gotree UpdateTipIndex: tip index = rank of the label in ascending byte order; duplicate
  // labels are an error (reported by the callers as ErrMulTree, preprocess.go:31-33,87-89).
  bool update_tip_index() {
    std::vector<int> t = tips();
    std::vector<std::pair<std::string, int>> byname;
    byname.reserve(t.size());
    for (int x : t) byname.emplace_back(nodes[x].name, x);
    std::sort(byname.begin(), byname.end());
    for (size_t i = 0; i + 1 < byname.size(); ++i)
      if (byname[i].first == byname[i + 1].first) return false;
    for (size_t i = 0; i < byname.size(); ++i) nodes[byname[i].second].tip_index = (int)i;
    return true;
  }

  // gotree RemoveSingleNodes: splice out every non-root node with exactly one child.
  void remove_single_nodes() {
    for (int x : postorder()) {
      Node& nd = nodes[x];
      if (nd.parent >= 0 && nd.children.size() == 1) {
        int p = nd.parent, c = nd.children[0];
        auto& pc = nodes[p].children;
        *std::find(pc.begin(), pc.end(), x) = c;
        nodes[c].parent = p;
        nd.children.clear();
        nd.parent = -2;  // detached
      }
    }
  }

  // gotree CollapseLowSupport(t, true): contract every internal (non-tip) edge whose
  // support is set and < t, root-adjacent ones included. Children take the contracted node's
  // place in its parent's child list.
  void collapse_low_support(double thr) {
    for (int x : postorder()) {
      Node& nd = nodes[x];
      if (nd.parent >= 0 && !nd.tip() && nd.has_support && nd.support < thr) {
        int p = nd.parent;
        auto& pc = nodes[p].children;
        auto it = std::find(pc.begin(), pc.end(), x);
        size_t pos = it - pc.begin();
        std::vector<int> kids = nd.children;
        pc.erase(pc.begin() + pos);
        pc.insert(pc.begin() + pos, kids.begin(), kids.end());
        for (int c : kids) nodes[c].parent = p;
        nd.children.clear();
        nd.parent = -2;
      }
    }
  }

  void clear_supports() {
    for (auto& n : nodes) {
      n.has_support = false;
      n.support = 0.0;
    }
  }

  // Newick string with names only (gotree Newick() after cleanTree, graphs/net.go:93-101).
  std::string newick() const {
    std::string out;
    write_newick(root, out);
    out += ';';
    return out;
  }

 private:
  std::vector<int> preorder_mirrored() const {  // root, children right-to-left
    std::vector<int> out, stack;
    if (root < 0) return out;
    stack.push_back(root);
    while (!stack.empty()) {
      int x = stack.back();
      stack.pop_back();
      out.push_back(x);
      for (int c : nodes[x].children) stack.push_back(c);
    }
    return out;
  }
  void write_newick(int x, std::string& out) const {
    const Node& nd = nodes[x];
    if (!nd.children.empty()) {
      out += '(';
      for (size_t i = 0; i < nd.children.size(); ++i) {
        if (i) out += ',';
        write_newick(nd.children[i], out);
      }
      out += ')';
    }
    out += nd.name;
  }
};

// ---------------------------------------------------------------------------------------------
// Newick reader: a label after ')' that parses as a number is the support of the edge
// above that node, otherwise the node name; '[...]' comments are skipped; ':len' is read and
// dropped; single-quoted labels are accepted.
// ---------------------------------------------------------------------------------------------
class NewickParser {
 public:
  explicit NewickParser(const std::string& s) : s_(s) {}

  Tree parse() {
    Tree t;
    skip_ws_comments();
    if (pos_ >= s_.size()) throw CamusError(Err::InvalidFormat, "empty newick string");
    if (s_[pos_] != '(') throw CamusError(Err::InvalidFormat, "newick must start with '('");
    t.root = parse_subtree(t, -1);
    skip_ws_comments();
    if (pos_ >= s_.size() || s_[pos_] != ';')
      throw CamusError(Err::InvalidFormat, "newick does not end with ';'");
    ++pos_;
    return t;
  }

 private:
  const std::string& s_;
  size_t pos_ = 0;

  void skip_ws_comments() {
    while (pos_ < s_.size()) {
      char c = s_[pos_];
      if (c == ' ' || c == '\t' || c == '\n' || c == '\r') {
        ++pos_;
      } else if (c == '[') {
        size_t e = s_.find(']', pos_);
        if (e == std::string::npos) throw CamusError(Err::InvalidFormat, "unterminated comment");
        pos_ = e + 1;
      } else {
        break;
      }
    }
  }

  std::string read_label() {
    skip_ws_comments();
    std::string lab;
    if (pos_ < s_.size() && s_[pos_] == '\'') {
      ++pos_;
      while (pos_ < s_.size()) {
        if (s_[pos_] == '\'') {
          if (pos_ + 1 < s_.size() && s_[pos_ + 1] == '\'') {
            lab += '\'';
            pos_ += 2;
            continue;
          }
          ++pos_;
          break;
        }
        lab += s_[pos_++];
      }
      return lab;
    }
    while (pos_ < s_.size()) {
      char c = s_[pos_];
      if (c == ',' || c == '(' || c == ')' || c == ':' || c == ';' || c == '[') break;
      lab += c;
      ++pos_;
    }
    while (!lab.empty() && (lab.back() == ' ' || lab.back() == '\t' || lab.back() == '\n' || lab.back() == '\r'))
      lab.pop_back();
    return lab;
  }

  void read_length() {
    skip_ws_comments();
    if (pos_ < s_.size() && s_[pos_] == ':') {
      ++pos_;
      skip_ws_comments();
      while (pos_ < s_.size()) {
        char c = s_[pos_];
        if (c == ',' || c == '(' || c == ')' || c == ';' || c == '[') break;
        ++pos_;
      }
      skip_ws_comments();
    }
  }

  static bool parse_number(const std::string& s, double& v) {
    if (s.empty()) return false;
    char* end = nullptr;
    v = std::strtod(s.c_str(), &end);
    return end && *end == '\0';
  }

  int parse_subtree(Tree& t, int parent) {
    // iterative descent would be safer for very deep trees; recursion depth here is the tree
    // height, which for 32767-taxon caterpillars is still within default stacks (tiny frames)
    int me = t.add_node();
    t.nodes[me].parent = parent;
    skip_ws_comments();
    if (pos_ < s_.size() && s_[pos_] == '(') {
      ++pos_;
      for (;;) {
        int c = parse_subtree(t, me);
        t.nodes[me].children.push_back(c);
        skip_ws_comments();
        if (pos_ >= s_.size()) throw CamusError(Err::InvalidFormat, "unbalanced parentheses");
        if (s_[pos_] == ',') {
          ++pos_;
          continue;
        }
        if (s_[pos_] == ')') {
          ++pos_;
          break;
        }
        throw CamusError(Err::InvalidFormat, "unexpected character in newick");
      }
      std::string lab = read_label();
      double v;
      if (parse_number(lab, v)) {
        t.nodes[me].has_support = true;
        t.nodes[me].support = v;
      } else {
        t.nodes[me].name = lab;
      }
    } else {
      t.nodes[me].name = read_label();
    }
    read_length();
    return me;
  }
};

inline Tree parse_newick(const std::string& s) { return NewickParser(s).parse(); }

inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (unsigned char)s[a] <= ' ') ++a;
  while (b > a && (unsigned char)s[b - 1] <= ' ') --b;
  return s.substr(a, b - a);
}

// prep.readTreeFile (io.go:99-119): exactly one newick tree; lengths, comments, supports cleared.
inline Tree read_constraint_text(const std::string& text, const std::string& fname = "<text>") {
  std::string body = trim(text);
  if (body.empty() || body.find('\n') != std::string::npos)
    throw CamusError(Err::InvalidFile,
                     "invalid file, there should only be exactly one newick tree in tree file " + fname);
  Tree t;
  try {
    t = parse_newick(body);
  } catch (const CamusError& e) {
    throw CamusError(Err::InvalidFormat, "invalid format, error parsing tree newick string from " +
                                             fname + ": " + e.what());
  }
  t.clear_supports();
  return t;
}

struct GeneTrees {
  std::vector<Tree> trees;
  std::vector<std::string> names;
};

// prep.readGeneTreesFile newick branch (io.go:133-153): one tree per non-blank line.
inline GeneTrees read_gene_trees_newick(const std::string& text, const std::string& fname = "<text>") {
  GeneTrees g;
  // non-blank lines first, then the trees are parsed in parallel (the reference parses serially;
  // the first bad line in file order is still the one reported)
  struct Line {
    size_t begin, end;
    int lineno;
  };
  std::vector<Line> lines;
  size_t start = 0;
  int lineno = 0;
  while (start <= text.size()) {
    size_t nl = text.find('\n', start);
    if (nl == std::string::npos) nl = text.size();
    size_t b = start, e = nl;
    while (b < e && std::isspace((unsigned char)text[b])) ++b;
    while (e > b && std::isspace((unsigned char)text[e - 1])) --e;
    if (e > b) lines.push_back({b, e, lineno});
    ++lineno;
    if (nl == text.size()) break;
    start = nl + 1;
  }
  if (lines.empty()) throw CamusError(Err::InvalidFile, "invalid file, empty gene tree file " + fname);
  g.trees.resize(lines.size());
  std::vector<std::string> errs(lines.size());
  std::vector<char> bad(lines.size(), 0);
  auto work = [&](size_t i) {
    try {
      g.trees[i] = parse_newick(text.substr(lines[i].begin, lines[i].end - lines[i].begin));
    } catch (const CamusError& e) {
      bad[i] = 1;
      errs[i] = e.what();
    }
  };
  const size_t nthreads = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), text.size() > (1u << 18) ? lines.size() : 1);
  if (nthreads <= 1) {
    for (size_t i = 0; i < lines.size(); ++i) work(i);
  } else {
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (size_t t = 0; t < nthreads; ++t)
      th.emplace_back([&] {
        for (size_t i; (i = next.fetch_add(1)) < lines.size();) work(i);
      });
    for (auto& t : th) t.join();
  }
  for (size_t i = 0; i < lines.size(); ++i)
    if (bad[i])
      throw CamusError(Err::InvalidFormat, "invalid format, error reading gene tree on line " + std::to_string(lines[i].lineno) +
                                               " in " + fname + ": " + errs[i]);
  for (size_t i = 0; i < g.trees.size(); ++i) g.names.push_back(std::to_string(i + 1));
  return g;
}

// prep.readGeneTreesFile nexus branch (io.go:154-165): TREES block, optional TRANSLATE table.
inline GeneTrees read_gene_trees_nexus(const std::string& text, const std::string& fname = "<text>") {
  auto upper = [](std::string s) {
    for (auto& c : s) c = (char)std::toupper((unsigned char)c);
    return s;
  };
  GeneTrees g;
  std::string head = upper(trim(text).substr(0, 6));
  if (head != "#NEXUS") throw CamusError(Err::InvalidFormat, "invalid format, error reading gene tree nexus file " + fname + ": missing #NEXUS");
  // strip comments outside of newick strings is unsafe ([&R] etc.), so split on ';' first
  std::vector<std::string> stmts;
  {
    std::string cur;
    int depth = 0;  // bracket comment depth
    bool quoted = false;
    for (char c : text) {
      if (quoted) {
        cur += c;
        if (c == '\'') quoted = false;
        continue;
      }
      if (depth == 0 && c == '\'') {
        quoted = true;
        cur += c;
        continue;
      }
      if (c == '[') ++depth;
      if (c == ']' && depth > 0) {
        --depth;
        cur += c;
        continue;
      }
      if (c == ';' && depth == 0) {
        stmts.push_back(trim(cur));
        cur.clear();
        continue;
      }
      cur += c;
    }
  }
  bool in_trees = false;
  std::unordered_map<std::string, std::string> translate;
  for (auto& st : stmts) {
    std::string u = upper(st);
    if (u.rfind("#NEXUS", 0) == 0) {
      size_t p = u.find("BEGIN");
      if (p == std::string::npos) continue;
      st = trim(st.substr(p));
      u = upper(st);
    }
    if (u.rfind("BEGIN", 0) == 0) {
      in_trees = upper(trim(st.substr(5))) == "TREES";
      continue;
    }
    if (u == "END" || u == "ENDBLOCK") {
      in_trees = false;
      continue;
    }
    if (!in_trees) continue;
    if (u.rfind("TRANSLATE", 0) == 0) {
      std::string body = st.substr(9);
      size_t s0 = 0;
      while (s0 < body.size()) {
        size_t c = body.find(',', s0);
        if (c == std::string::npos) c = body.size();
        std::string item = trim(body.substr(s0, c - s0));
        size_t sp = item.find_first_of(" \t\n\r");
        if (sp != std::string::npos) {
          std::string key = item.substr(0, sp), val = trim(item.substr(sp));
          if (val.size() >= 2 && val.front() == '\'' && val.back() == '\'') val = val.substr(1, val.size() - 2);
          translate[key] = val;
        }
        s0 = c + 1;
      }
      continue;
    }
    if (u.rfind("TREE", 0) == 0 && u.size() > 4 && (u[4] == ' ' || u[4] == '\t' || u[4] == '\n')) {
      size_t eq = st.find('=');
      if (eq == std::string::npos) throw CamusError(Err::InvalidFormat, "invalid format, error reading gene tree nexus file " + fname + ": tree statement without '='");
      std::string name = trim(st.substr(4, eq - 4));
      if (!name.empty() && name[0] == '*') name = trim(name.substr(1));
      std::string nwk = trim(st.substr(eq + 1)) + ";";
      Tree t;
      try {
        t = parse_newick(nwk);
      } catch (const CamusError& e) {
        throw CamusError(Err::InvalidFormat, "invalid format, error reading gene tree nexus file " + fname + ": " + e.what());
      }
      if (!translate.empty())
        for (auto& nd : t.nodes)
          if (nd.tip()) {
            auto it = translate.find(nd.name);
            if (it != translate.end()) nd.name = it->second;
          }
      g.trees.push_back(std::move(t));
      g.names.push_back(name);
    }
  }
  return g;
}

}  // namespace camus
