// `camus` command line, same flags, positional arguments, log lines and CSV output as the
// reference's camus.go (:82-87 defaults, :142-197 parseArgs, :219-283 main/run), with the hot path
// running on the GPU through libcamus_b200.so. Not reproduced: the PNG line plot
// (internal/prep/io.go:271-312, gonum); CAMUS_B200_DEVICE selects the CUDA ordinal (no new flag).
#include <charconv>
#include <chrono>
#include <cstdio>
#include <ctime>
#include <fstream>
#include <unistd.h>
#include <future>
#include <iostream>
#include <sstream>

#include "infer.hpp"

using namespace camus;

namespace {

const char* kVersion = "camus-b200 0.1.0";
const char* kErrorMessage = "camus encountered an error ::";

std::string g_logbuf;
FILE* g_logfile = nullptr;

void logline(const std::string& msg) {  // log.LstdFlags | log.Lmicroseconds
  auto now = std::chrono::system_clock::now();
  std::time_t t = std::chrono::system_clock::to_time_t(now);
  auto us = std::chrono::duration_cast<std::chrono::microseconds>(now.time_since_epoch()).count() % 1000000;
  std::tm tm{};
  localtime_r(&t, &tm);
  char head[64];
  snprintf(head, sizeof head, "%04d/%02d/%02d %02d:%02d:%02d.%06ld ", tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday,
           tm.tm_hour, tm.tm_min, tm.tm_sec, (long)us);
  std::string line = std::string(head) + msg + "\n";
  fputs(line.c_str(), stderr);
  if (g_logfile) {
    fputs(line.c_str(), g_logfile);
    fflush(g_logfile);
  } else {
    g_logbuf += line;
  }
}

void usage(bool extended) {
  fputs("usage: camus [flags]... <const_tree_file> <gene_tree_file>\n\npositional arguments:\n\n"
        "  <tree_file>\t\tconstraint newick tree\n  <gene_tree_file>\tgene tree newick file\n\nflags:\n\n",
        stderr);
  if (extended) {
    fputs("  -a float\n    \tparameter to adjust penalty for \"sym\" score mode, from (0, 1] (default 0.1)\n"
          "  -asSet\n    \tquartet count is calculated as a set (one point per unique topology)\n", stderr);
  }
  fputs("  -f format\n    \tgene tree format [newick|nexus] (default \"newick\")\n"
        "  -h\tprints short help and exits\n  -hh\n    \tprints help with experimental features and exits\n"
        "  -n int\n    \tnumber of parallel processes\n  -o string\n    \toutput prefix\n", stderr);
  if (extended)
    fputs("  -q int\n    \tquartet filter mode number [0, 2] (default 2)\n", stderr);
  fputs("  -s float\n    \tcollapse edges in gene trees with support less than value (default 0)\n", stderr);
  if (extended) fputs("  -sm mode\n    \tscore mode [max|norm|sym] (default \"max\")\n", stderr);
  fputs("  -t float\n    \tthreshold for quartet filter [0, 1] (default 0.5)\n  -v\tprints version number and exits\n"
        "\nexamples:\n\n\tcamus -o output-name constraint.nwk gene-trees.nwk\n\n", stderr);
}

[[noreturn]] void parser_error(const std::string& m) {
  fprintf(stderr, "%s\n\n", m.c_str());
  usage(false);
  exit(1);
}

std::string read_file(const std::string& path, const char* what) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw CamusError(Err::InvalidFile, std::string("error ") + what + " " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

std::string parse_name(std::string s) {  // camus.go:206-214
  size_t sl = s.find_last_of('/');
  if (sl != std::string::npos) s = s.substr(sl + 1);
  size_t dot = s.find_last_of('.');
  if (dot != std::string::npos && dot > 0) s = s.substr(0, dot);
  return s;
}

std::string format_float(double v) {  // strconv.FormatFloat(v, 'f', -1, 64)
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v > 0 ? "+Inf" : "-Inf";
  char buf[512];
  auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
  return std::string(buf, r.ptr);
}

std::string csv_field(const std::string& s) {  // encoding/csv quoting rules
  if (s.empty()) return s;
  bool q = s.find_first_of(",\"\r\n") != std::string::npos || s[0] == ' ';
  if (!q) return s;
  std::string o = "\"";
  for (char c : s) {
    if (c == '"') o += '"';
    o += c;
  }
  return o + "\"";
}

void write_csv(std::ostream& os, const TreeData& td, const std::vector<std::string>& newicks, const std::vector<double>& qsat) {
  os << "Number of Branches,Quartet Satisfied Percent,Extended Newick\n";  // io.go:241-269
  os << "0,0," << csv_field(td.tree.newick()) << "\n";
  for (size_t i = 0; i < newicks.size(); ++i) os << (i + 1) << "," << format_float(qsat[i]) << "," << csv_field(newicks[i]) << "\n";
}

}  // namespace

// Go's flag package rejects trailing garbage ("-q 2x"); stoi / stod alone would accept it
static int whole_int(const std::string& v) {
  size_t pos = 0;
  int r = std::stoi(v, &pos);
  if (pos != v.size()) throw std::invalid_argument(v);
  return r;
}
static double whole_double(const std::string& v) {
  size_t pos = 0;
  double r = std::stod(v, &pos);
  if (pos != v.size()) throw std::invalid_argument(v);
  return r;
}

int main(int argc, char** argv) {
  std::string format = "newick", prefix, scoreMode = "max";
  int qmode = 2, nprocs = 0;
  double supp = 0, thresh = 0.5, alpha = 0.1;
  bool asSet = false;
  std::vector<std::string> pos;
  std::string invoked;
  for (int i = 1; i < argc; ++i) invoked += (i > 1 ? " " : "") + std::string(argv[i]);
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (!pos.empty() || a.size() < 2 || a[0] != '-') {  // Go's flag package stops at the first non-flag
      pos.push_back(a);
      continue;
    }
    std::string name = a.substr(a[1] == '-' ? 2 : 1), val;
    bool has_val = false;
    size_t eq = name.find('=');
    if (eq != std::string::npos) {
      val = name.substr(eq + 1);
      name = name.substr(0, eq);
      has_val = true;
    }
    auto need = [&]() -> std::string {
      if (has_val) return val;
      if (i + 1 >= argc) parser_error("flag needs an argument: -" + name);
      return argv[++i];
    };
    try {
      if (name == "h") {
        usage(false);
        return 0;
      } else if (name == "hh") {
        usage(true);
        return 0;
      } else if (name == "v") {
        puts(kVersion);
        return 0;
      } else if (name == "asSet") {
        asSet = !has_val || val == "true" || val == "1";
      } else if (name == "f") {
        format = need();
        if (format != "newick" && format != "nexus") parser_error("invalid value \"" + format + "\" for flag -f: \"" + format + "\" is not a valid gene tree file format");
      } else if (name == "o") prefix = need();
      else if (name == "sm") scoreMode = need();
      else if (name == "q") qmode = whole_int(need());
      else if (name == "s") supp = whole_double(need());
      else if (name == "t") thresh = whole_double(need());
      else if (name == "a") alpha = whole_double(need());
      else if (name == "n") nprocs = whole_int(need());
      else parser_error("flag provided but not defined: -" + name);
    } catch (const std::invalid_argument&) {
      parser_error("invalid value for flag -" + name);
    } catch (const std::out_of_range&) {  // Go's flag package: "value out of range"
      parser_error("invalid value for flag -" + name + ": value out of range");
    }
  }
  if (pos.size() != 2) parser_error("two positional arguments required: <const_tree> <gene_tree_file>");
  ScoreMode mode;
  if (!ParseScorer(scoreMode, mode))
    parser_error("\"" + scoreMode + "\" is not a valid score mode: valid score modes are \"max\", \"norm\", and \"sym\"");
  QuartetFilterOptions qopts;
  try {
    qopts = SetQuartetFilterOptions(qmode, thresh);
  } catch (const CamusError& e) {
    parser_error(e.what());
  }
  if (prefix.empty()) {
    std::time_t t = std::time(nullptr);
    std::tm tm{};
    localtime_r(&t, &tm);
    char ts[32];
    strftime(ts, sizeof ts, "%Y-%m-%d_%H-%M-%S", &tm);
    prefix = "camus_" + parse_name(pos[0]) + "_" + parse_name(pos[1]) + "_" + ts;
    logline("output prefix was not set, using \"" + prefix + "\"");
  }
  g_logfile = fopen((prefix + ".log").c_str(), "w");
  if (g_logfile) {
    fputs(g_logbuf.c_str(), g_logfile);
  } else {
    logline("failed to create log file " + prefix + ".log");
  }
  logline(std::string("camus ") + kVersion);
  logline("invoked as: camus " + invoked);
  int exit_code = 0;
  try {
    InferOptions opts = MakeInferOptions(nprocs, qopts, supp, mode, asSet, alpha, logline);
    // CAMUS_B200_DEVICE = first CUDA ordinal (default 0), CAMUS_B200_GPUS = how many GPUs from there on
    // (default 1). The reference CLI has no such flags; the environment keeps its flag set unchanged.
    int ordinal = getenv("CAMUS_B200_DEVICE") ? atoi(getenv("CAMUS_B200_DEVICE")) : 0;
    int ngpus = getenv("CAMUS_B200_GPUS") ? std::max(1, atoi(getenv("CAMUS_B200_GPUS"))) : 1;
    std::vector<int> ordinals;
    for (int i = 0; i < ngpus; ++i) ordinals.push_back(ordinal + i);
    // creating the CUDA context(s) costs about a second: it runs while the input files are parsed
    auto dev_ready = std::async(std::launch::async, [ordinals] { return std::make_shared<Device>(ordinals); });
    Tree tre = read_constraint_text(read_file(pos[0], "reading tree file:"), pos[0]);
    std::string gtext = read_file(pos[1], "opening");
    GeneTrees gts = format == "nexus" ? read_gene_trees_nexus(gtext, pos[1]) : read_gene_trees_newick(gtext, pos[1]);
    auto dev = dev_ready.get();
    InferResults res = Infer(std::move(tre), gts.trees, opts, dev, logline);
    std::vector<std::string> newicks;
    for (auto& b : res.dp.branches) newicks.push_back(make_network_newick(res.tree.td, b));
    write_csv(std::cout, res.tree.td, newicks, res.dp.qsat);
    std::ofstream f(prefix + ".csv");
    if (!f) throw CamusError(Err::InvalidFile, "cannot create " + prefix + ".csv");
    write_csv(f, res.tree.td, newicks, res.dp.qsat);
    logline("line plot (" + prefix + ".png) is not produced by this build");
  } catch (const std::exception& e) {
    logline(std::string(kErrorMessage) + " " + e.what());
    exit_code = 1;
  }
  if (g_logfile) fclose(g_logfile);
  // every output is written: leave without releasing tens of GB of device memory piece by piece
  // and tearing the CUDA context down (0.3 - 0.6 s); the driver reclaims both
  std::cout.flush();
  fflush(nullptr);
  _exit(exit_code);
}
