"""Known-answer tests transcribed BY HAND from the reference's own Go tests (values only, cited
file:line relative to /root/reference). They pin the oracle (tests/test_oracle_golden.py) and,
through it, the CUDA path. Data files copied verbatim from the reference's testdata/ directories
live in tests/golden/ref/ (see tests/golden/README.md).

A quartet written ((a,b),(c,d)) is given as the tuple ("a","b","c","d") = pairing ab|cd.
"""

# internal/prep/preprocess_test.go:130-224  (TestProcessQuartets)
# (name, constraint, qmode, threshold, gene trees, expected surviving quartets with multiplicity)
PROCESS_QUARTETS = [
    ("basic", "((((a,b),c),d),f);", 0, 0.0,
     ["(((a,b),c),d);", "(((a,b),c),f);", "(((a,b),d),f);", "(((c,d),f),a);", "(((d,b),a),f);", "((c,f),(d,b));"],
     [("c", "d", "f", "a"), ("b", "d", "a", "f"), ("c", "f", "d", "b")]),
    ("q mode 1", "((((a,b),c),d),f);", 1, 0.0,
     ["(((a,b),c),d);", "(((a,b),c),f);", "(((a,b),d),f);", "(((c,d),f),a);", "(((d,b),a),f);", "((c,f),(d,b));",
      "((c,b),(d,f));", "((c,b),(d,f));", "((c,d),(f,b));", "((c,d),(f,b));"],
     [("c", "d", "f", "a"), ("b", "d", "a", "f"), ("c", "d", "f", "b"), ("c", "d", "f", "b"), ("c", "f", "d", "b")]),
    ("q mode 2", "((((a,b),c),d),f);", 2, 0.0,
     ["(((a,b),c),d);", "(((a,b),c),f);", "(((a,b),d),f);", "(((c,d),f),a);", "(((d,b),a),f);", "((c,f),(d,b));",
      "((c,d),(f,b));", "((c,d),(f,b));"],
     [("c", "f", "d", "b"), ("c", "d", "f", "a"), ("b", "d", "a", "f"), ("c", "d", "f", "b"), ("c", "d", "f", "b")]),
    ("unresolved gene tree simple", "((((a,b),c),d),f);", 0, 0.0,
     ["((a,c,d),(b,f));"],
     [("a", "c", "b", "f"), ("a", "d", "b", "f"), ("c", "d", "b", "f")]),
    ("unresolved gene tree", "((((a,b),c),d),f);", 0, 0.0,
     ["(a,c,d,b,f);"],
     []),
]

# internal/graphs/quartet_test.go:135-180  (TestQuartetsFromTree)
QUARTETS_FROM_TREE = [
    ("basic", "((((a,b),c),d),f);",
     [("a", "b", "c", "d"), ("a", "b", "c", "f"), ("a", "b", "d", "f"), ("a", "c", "d", "f"), ("b", "c", "d", "f")]),
    ("single", "((a,b),(c,d));", [("a", "b", "c", "d")]),
]

# internal/graphs/quartet_test.go:19-68 (TestNewQuartet): both trees give the quartet ab|cd on the
# constraint (((a,c),(b,d)),f); (tip indices a0 b1 c2 d3 f4). String() (quartet.go:179-199) puts
# taxon i left of '|' when bit i of the topology is set.
NEW_QUARTET = [("a", "b", "c", "d"), ]

# internal/graphs/quartet_test.go:70-133 (TestCompare): results Qeq / Qneq / Qdiff
COMPARE = [
    (("a", "b", "c", "d"), ("a", "b", "c", "d"), "eq"),
    (("a", "b", "c", "d"), ("a", "c", "b", "d"), "neq"),
    (("a", "b", "c", "d"), ("a", "f", "b", "d"), "diff"),
]

# internal/score/quartettotals_test.go:71-119
SHOULD_CALC_TREE = "((A,(B,C)b)a,(D,E)c)r;"
SHOULD_CALC_EDGE = [("b", "c", True), ("b", "a", False), ("a", "c", False), ("r", "b", False)]
CYCLE_LENGTH = [("b", "c", 4), ("a", "c", 3), ("A", "E", 5), ("b", "B", 3)]

# internal/score/quartettotals_test.go:121-153 (TestQuartetsTotal), :12-69 (table == per-pair)
TOTALS_TREE_A = ("((A,B)a,(C,D)b)r;", [(("A", "C", "B", "D"), 5)])
TOTALS_TREE_B = ("(((A,B)a,(C,D)b)e,(E,(F,G)f)c)r;", [(("A", "E", "B", "F"), 7), (("A", "F", "B", "E"), 4)])
QUARTETS_TOTAL = [
    ("A", TOTALS_TREE_A, "A", "C", 5), ("A", TOTALS_TREE_A, "C", "A", 5), ("A", TOTALS_TREE_A, "A", "B", 0),
    ("B", TOTALS_TREE_B, "A", "E", 7), ("B", TOTALS_TREE_B, "E", "A", 7), ("B", TOTALS_TREE_B, "A", "F", 4),
    ("B", TOTALS_TREE_B, "B", "C", 0),
]

# internal/score/quartettotals_test.go:155-230 (BenchmarkQuartetScore pre-checks)
QUARTET_SCORE = [
    ("((A,B)a,(C,D)b)r;", ("A", "C", "B", "D"), "A", "C", "eq"),
    ("((A,B)a,(C,D)b)r;", ("A", "D", "B", "C"), "A", "C", "neq"),
    ("((A,B)a,(C,D)b)r;", ("A", "C", "B", "D"), "A", "B", "diff"),
]

# internal/score/penalty_test.go:25-116
PENALTY_TREE = "((((A,B)a,C)b,D)c,E)r;"
PENALTIES = [("b", "a", 4), ("c", "a", 8)]

# internal/graphs/treedata_test.go:12-95 (TestMakeTreeData): LCAs, and quartet sets per node
TREEDATA_TREE = "((((A,B)a,C)b,D)c,F)r;"
TREEDATA_LCA = {"a": [("A", "B")], "b": [("A", "C"), ("B", "C")], "c": [("A", "D"), ("B", "D"), ("C", "D")],
                "r": [("A", "F"), ("B", "F"), ("C", "F"), ("D", "F")]}
TREEDATA_QSETS = (("A", "C", "B", "D"), {"a": 0, "b": 0, "c": 1, "r": 1})

# internal/infer/infer_test.go:17-174 (TestInfer): -q 0, max mode; number of edges and the
# network with all edges.
INFER = [
    ("basic one-edge", "(A,(B,(C,(D,(E,(F,(G,(H,(I,J)))))))));",
     ["(A,(B,(C,D)));", "(B,(C,D),E);"], 1,
     "(A,(B,((C)#H1,((#H1,D),(E,(F,(G,(H,(I,J)))))))));"),
    ("basic two-edges", "((A,((((B,C),D),E),F)),(G,H));",
     ["((A,B),(C,D));", "((G,F),(A,H));"], 2,
     "(((A)#H1,((((B,(C)#H2),(#H2,D)),E),F)),(G,(#H1,H)));"),
    ("two-edge two", "(A,(B,(C,(D,(E,(F,(G,(H,(I,J)))))))));",
     ["((J,G),(H,I));", "((C,G),(E,F));"], 2,
     "(A,(B,(C,(D,((E)#H1,((#H1,F),(G,((H)#H2,((#H2,I),J)))))))));"),
    ("two-edge case two", "((A,((((B,C),D),E),F)),(G,H));",
     ["((A,B),(C,D));", "((A,F),(G,E));"], 2,
     "(((A)#H1,((((B,(C)#H2),(#H2,D)),E),(#H1,F))),(G,H));"),
    ("one-sided cycle test", "((((A,B),C),D),E);",
     ["((B,C),(D,A));", "((B,C),(D,E));", "((B,C),(A,E));", "((B,D),(A,E));", "((C,D),(A,E));"], 1,
     "((#H1,((((A)#H1,B),C),D)),E);"),
    ("double one-sided cycle test", "(((((((((A,B),C),D),E),F),G),H),I),J);",
     ["((B,C),(D,A));", "((B,C),(D,E));", "((B,C),(A,E));", "((B,D),(A,E));", "((C,D),(A,E));",
      "((G,H),(I,D));", "((G,H),(I,J));", "((G,H),(D,J));", "((G,I),(D,J));", "((H,I),(D,J));", "((E,I),(A,J));"], 2,
     "((#H1,(((((((#H2,((((A)#H2,B),C),D)))#H1,E),F),G),H),I)),J);"),
    ("duplicate quartet basic", "(((((A,B),C),D),E),F);",
     ["((E,D),(F,C));", "((C,B),(A,D));", "((D,C),(A,E));", "((D,C),(A,E));", "((D,C),(A,E));"], 1,
     "(((#H1,((((A)#H1,B),C),D)),E),F);"),
    ("duplicate quartet basic 2", "(((((A,B),C),D),E),F);",
     ["((E,C),(F,B));", "((E,C),(F,B));", "((E,C),(F,B));", "((C,B),(A,D));", "((D,C),(A,E));"], 1,
     "((#H1,(((((A,B))#H1,C),D),E)),F);"),
    ("avoid over-adding edges", "(R,((A,(((B,C),D),((E,F),G))),H));",
     ["((C,D),(B,H));", "((F,G),(E,H));", "((R,A),(B,H));"], 2,
     "(R,((A,((((B)#H1,C),D),((E,(F)#H2),(#H2,G)))),(#H1,H)));"),
    ("avoid over-adding edges 2", "(R,((A,(((B,C),D),((E,F),G))),H));",
     ["((C,D),(B,H));", "((F,G),(E,H));", "((R,A),(B,H));", "((R,D),(E,H));"], 2,
     "(R,((A,(((B,(C)#H2),(#H2,D)),(((#H1,E),F),G))),(H)#H1));"),
    ("test under node u lookup", "(R,((A,(I,J)),((((B,C),D),H),((E,F),G))));",
     ["((C,D),(B,A));", "((F,G),(E,A));", "((H,A),(E,C));", "((H,R),(E,C));", "((I,R),(J,A));"], 3,
     "(R,(((A)#H1,(I,(#H1,J))),(((#H2,((B,(C)#H3),(#H3,D))),H),(((E)#H2,F),G))));"),
    ("cycle below base of one-sided cycle", "(((((A,B),C),(D,(F,(G,H)))),E),R);",
     ["((B,C),(D,A));", "((B,C),(D,E));", "((B,C),(A,E));", "((B,D),(A,E));", "((C,D),(A,E));",
      "((F,G),(H,E));", "((F,G),(H,E));", "((E,C),(A,R));"], 2,
     "((#H1,(((((A)#H1,B),C),(D,((F)#H2,((#H2,G),H)))),E)),R);"),
]

# internal/infer/infer_test.go:208-301 (TestInfer_Large): (qmode, threshold, score mode, alpha,
# expected number of networks, golden file under tests/golden/ref/)
INFER_LARGE = [
    (0, 0.0, "max", 0.0, 5, "network.nwk"),
    (2, 0.5, "max", 0.0, 4, "net_q2_t05_max.nwk"),
    (2, 0.5, "norm", 0.0, 4, "net_q2_t05_norm.nwk"),
    (2, 0.5, "sym", 0.1, 4, "net_q2_t05_sym_a01.nwk"),
]

# internal/prep/preprocess_test.go:16-128 (TestIsBinary / TestPreprocess_Errors)
PREPROCESS_ERRORS = [
    ("unrooted", "((a,b),(c,d),e);", [], "ErrUnrooted"),
    ("non-binary", "((a,b),(c,d,e));", [], "ErrNonBinary"),
    ("multree", "((a,a),(a,a));", [], "ErrMulTree"),
    ("multree gtree", "((a,b),(c,d));", ["((a,a),(a,a));"], "ErrMulTree"),
    ("missing const labels", "((a,b),(c,d));", ["((a,b),(c,d));", "(((a,b),(c,d)),e);"], "ErrTipNameMismatch"),
    ("non-binary input trees", "((a,b),(c,d));", ["(a,b,c,d);"], None),
    ("unifurcation success", "(((a,b)),(c,d));", [], None),
]
