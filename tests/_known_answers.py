"""The reference's own known-answer tests that run through the balancing path
(reference test/test_all.py:312-334 test_09_hic_normalization and :380-405
test_11_write_interaction_pairs), restated over a bias vector: what those tests assert is a
function of the raw cells, the bad columns and the biases only.

    test_09   no column filter; normalize_hic() defaults; sum over i < j of
              hic[i,j] / bias[i] / bias[j] over the non-zero cells, rounded to 4 decimals,
              is 4059.2877 (test_all.py:332)
    test_11   filter_columns(); normalize_hic(factor=1); the pair file has 4674 lines and
              lines 25 / 2000 hold 0.5852295196345679 / 0.07060448846960976 (test_all.py:400-402)

tests/golden/reference_known_answers.json holds the same quantities produced by the UNMODIFIED
reference classes (oracle/gen_golden.py:reference_known_answers)."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KNOWN = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))

# literals asserted by the reference's test file
TEST_09_SUM = 4059.2877
TEST_11_LINES = 4674
TEST_11_LINE_25 = 0.5852295196345679
TEST_11_LINE_2000 = 0.07060448846960976


def pair_lines(case, bias, bads):
    """(i, j, normalised contact) for i < j, both columns kept, cell non-zero, row-major:
    Experiment.normalize_hic (experiment.py:737-743) + get_hic_zscores(zscored=False) +
    write_interaction_pairs() (experiment.py:751-799, :1078-1190)."""
    cells = {}
    for r, c, v in zip(case.rows.tolist(), case.cols.tolist(), case.vals.tolist()):
        if r != c and v:
            cells[(min(r, c), max(r, c))] = v
    out = []
    for (i, j) in sorted(cells):
        if i in bads or j in bads:
            continue
        out.append((i, j, float(cells[(i, j)]) / bias[i] / bias[j]))
    return out
