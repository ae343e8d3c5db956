#!/usr/bin/env python
"""Extracts the known-answer literals of the reference's own unit tests into
small JSON fixtures (run in the build container, where /root/reference exists;
the fixtures are committed, the reference sources are not copied).

  bsmbsm_solve.json     tests/test_bsmbsm.cpp:698-798  M, D1, D2 (general band
                        storage, kl=ku=4, n=10), BR, XR (30-digit Mathematica
                        solution of the S=3 block system op x = BR)
  bsmbsm_perm.json      tests/test_bsmbsm.cpp:48-152   q / qinv for S=5, n=9
  bsplineop_colloc.json tests/test_bsplineop.cpp:70-742 Greville collocation
                        operators D_T[d] for k = 2, 3, 4 (and the k=4
                        single-interval "almost dense" case)
  lapack_gbtrf.json     tests/test_lapack.c:35-76       4x4 dgbtrf + dgbcon case
  rholut_known_answers.json tests/test_rholut.cpp:37-108, 155-240  the test field (rho, m, e at (1,2,3)) and the
                        200-bit Sage values of p, T, mu, lambda for alpha=5, beta=2/3, gamma=1.4, Ma=3.5

Usage: python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    return re.sub(r"//[^\n]*", " ", s)


def numbers(body):
    """Evaluates a brace-enclosed C initialiser list of arithmetic constants."""
    body = strip_comments(body)
    vals = []
    for tok in body.split(","):
        tok = tok.strip()
        if not tok:
            continue
        tok = re.sub(r"-\s+", "-", tok)                       # "- 5.2" -> "-5.2"
        tok = re.sub(r"(?<![\w.])(\d+)\.(?![\d])", r"\1.0", tok)  # "1." -> "1.0"
        if not re.fullmatch(r"[-+*/(). \deE]+", tok):
            raise ValueError(f"unexpected token {tok!r}")
        vals.append(float(eval(tok, {"__builtins__": {}})))
    return vals


def array(src, name, start=0):
    m = re.compile(r"(?:static\s+)?const\s+double\s+" + re.escape(name) + r"\s*\[\s*\]\s*=\s*\{(.*?)\}\s*;",
                   re.S).search(src, start)
    if not m:
        raise KeyError(name)
    return numbers(m.group(1)), m.end()


def bsmbsm():
    raw = open(os.path.join(REF, "tests/test_bsmbsm.cpp")).read()
    start = raw.index("BOOST_AUTO_TEST_SUITE( gbsv )") if "BOOST_AUTO_TEST_SUITE( gbsv )" in raw else 0
    out = {"source": "tests/test_bsmbsm.cpp", "S": 3, "n": 10, "kl": 4, "ku": 4,
           "op": "[[M, D1/5, 0], [D2/7, 2M, D2/14], [0, D1/10, 4M]]",
           "tolerance": "1.8e4 * eps relative (check_close_collections)"}
    for name in ("M", "D1", "D2", "BR", "XR"):
        out[name], _ = array(raw, name, start)
    assert len(out["M"]) == len(out["D1"]) == len(out["D2"]) == 90 and len(out["XR"]) == 30
    json.dump(out, open(os.path.join(OUT, "bsmbsm_solve.json"), "w"), indent=0)

    src = strip_comments(raw)
    perm = {"source": "tests/test_bsmbsm.cpp:48-152", "S": 5, "n": 9, "q": {}, "qinv": {}}
    for fn in ("q", "qinv"):
        for m in re.finditer(r"BOOST_CHECK_EQUAL\(\s*(\d+)\s*,\s*suzerain_bsmbsm_" + fn +
                             r"\(S,\s*n,\s*(\d+)\)\)", src):
            perm[fn][m.group(2)] = int(m.group(1))
    assert len(perm["q"]) == 45 and len(perm["qinv"]) == 45
    json.dump(perm, open(os.path.join(OUT, "bsmbsm_perm.json"), "w"), indent=0)


def bsplineop():
    raw = open(os.path.join(REF, "tests/test_bsplineop.cpp")).read()
    cases = []
    for m in re.finditer(r"BOOST_AUTO_TEST_CASE\(\s*(collocation_piecewise_\w+)\s*\)", raw):
        name = m.group(1)
        nxt = raw.find("BOOST_AUTO_TEST_CASE", m.end())
        body = raw[m.end(): nxt if nxt > 0 else len(raw)]
        bp, _ = array(body, "breakpts")
        k = int(re.search(r"bspline\s+b\(\s*(\d+)\s*,", body).group(1))
        case = {"name": name, "k": k, "breakpoints": bp, "D_T": []}
        for g in re.finditer(r"const\s+double\s+(good_D_T(\d))\s*\[\s*\]\s*=\s*\{(.*?)\}\s*;", body, re.S):
            chk = re.search(r"CHECK_GBMATRIX_CLOSE\(\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*(\d+)\s*,\s*"
                            + g.group(1) + r"\s*,\s*(\d+)\s*,", body[g.end():])
            if not chk:
                continue
            mm, nn, kl, ku, ld = map(int, chk.groups())
            vals = numbers(g.group(3))
            assert len(vals) == ld * nn, (name, g.group(1), len(vals), ld, nn)
            case["D_T"].append({"d": int(g.group(2)), "m": mm, "n": nn, "kl": kl, "ku": ku,
                                "ld": ld, "band": vals})
        if case["D_T"]:
            cases.append(case)
    assert len(cases) >= 4, [c["name"] for c in cases]
    # k = 8 on one interval: the reference checks operator actions only (:697-742)
    m = re.search(r"BOOST_AUTO_TEST_CASE\(\s*collocation_piecewise_septic_dense_and_then_some\s*\)", raw)
    body = raw[m.end(): raw.find("BOOST_AUTO_TEST_CASE", m.end())]
    septic = {"name": "collocation_piecewise_septic_dense_and_then_some", "k": 8,
              "breakpoints": array(body, "breakpts")[0], "actions": []}
    pos = 0
    while True:
        g = re.compile(r"double\s+b\s*\[\s*\]\s*=\s*\{(.*?)\}\s*;", re.S).search(body, pos)
        if not g:
            break
        ap = re.compile(r"op\.apply\(\s*(\d+)\s*,\s*(\d+|nrhs)\s*,\s*([-\d.eE]+)\s*,").search(body, g.end())
        good, pos = array(body, "b_good", g.end())
        septic["actions"].append({"d": int(ap.group(1)), "alpha": float(ap.group(3)), "nrhs": 2,
                                  "x": numbers(g.group(1)), "y": good,
                                  "tolerance": "1000 * eps relative"})
    assert len(septic["actions"]) == 2
    json.dump({"source": "tests/test_bsplineop.cpp:70-742", "cases": cases, "septic": septic},
              open(os.path.join(OUT, "bsplineop_colloc.json"), "w"), indent=0)
    return cases


def rholut():
    """tests/test_rholut.cpp: the test field of rholut_test_data() and the known answers of the
    rholut_p_T_mu_lambda case (long-double literals from test_rholut.sage)."""
    raw = strip_comments(open(os.path.join(REF, "tests/test_rholut.cpp")).read())
    num = r"(-?\s*[\d.]+(?:[eE][-+]?\d+)?)L?"
    data = raw[raw.index("void rholut_test_data("):]
    data = data[:data.index("BOOST_AUTO_TEST_CASE")]
    val = lambda name: float(re.search(re.escape(name) + r"\s*=\s*" + num + r"\s*;", data).group(1).replace(" ", ""))
    case = raw[raw.index("BOOST_AUTO_TEST_CASE( rholut_p_T_mu_lambda )"):]
    case = case[:case.index("BOOST_AUTO_TEST_CASE", 10)]
    par = lambda name: float(eval(re.search(r"const double " + name + r"\s*=\s*([^;]+);", case).group(1),
                                  {"__builtins__": {}}))
    ans = lambda name: float(re.search(r"BOOST_CHECK_CLOSE\(" + name + r",\s*" + num, case).group(1))
    out = {"source": "tests/test_rholut.cpp:37-108 (rholut_test_data), :155-240 (rholut_p_T_mu_lambda)",
           "tolerance": "1e3 * eps relative (BOOST_CHECK_CLOSE is in percent: 1e3 eps percent = 1e1 eps)",
           "rho": val("rho"), "m": [val("m(0)"), val("m(1)"), val("m(2)")], "e": val("e"),
           "alpha": par("alpha"), "beta": par("beta"), "gamma": par("gamma"), "Ma": par("Ma"),
           "p": ans("p"), "T": ans("T"), "mu": ans("mu"), "lambda": ans("lambda")}
    json.dump(out, open(os.path.join(OUT, "rholut_known_answers.json"), "w"), indent=1)
    return out


if __name__ == "__main__":
    bsmbsm()
    print("rholut:", rholut())
    cs = bsplineop()
    print("bsplineop cases:", [(c["name"], c["k"], [(d["d"], d["kl"], d["ku"]) for d in c["D_T"]]) for c in cs])
