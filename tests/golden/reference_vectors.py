"""Golden vectors transcribed (as DATA) from the reference's own Go tests.

Every table cites the reference test file:line it was transcribed from
(paths relative to the reference root, Metronlab/bow).  Only in-scope
(int64 / float64) vectors are kept; the Boolean/String variants of the same
tables are out of scope.

Go constant expressions such as `100*0.1 + 200*0.9` are evaluated by the Go
compiler in exact arithmetic before rounding to float64; the values below are
those exact results (e.g. 190.0), NOT the Python float expression.
"""

N = None  # null

TIME, VALUE = "time", "value"

# --------------------------------------------------------------------------
# rolling/rolling_test.go:20-68  TestIntervalRolling_NumWindows
# (cols, interval, offset, expected NumWindows)
# --------------------------------------------------------------------------
NUM_WINDOWS = [
    ("empty bow", [[], []], 1, 0, 0),
    ("one liner bow", [[0], [1.0]], 1, 0, 1),
    ("points in same window", [[0, 9], [1.0, 1.0]], 10, 0, 1),
    ("excluded point goes in next window", [[0, 10], [1.0, 1.0]], 10, 0, 2),
    ("offset puts first value in preceding window", [[0, 9], [1.0, 1.0]], 10, 1, 2),
]

# rolling/rolling_test.go:70-109  TestIntervalRolling_iterator_init
CTOR_ERRORS = [
    ("interval == 0", {"interval": 0}, "enforceIntervalAndOffset: strictly positive interval required"),
    ("non existing index", {"col": "badcol", "interval": 1}, "no column 'badcol'"),
    ("invalid interval type", {"time_type": "float64", "interval": 1},
     "impossible to create a new intervalRolling on column of type float64"),
]

# --------------------------------------------------------------------------
# rolling/rolling_test.go:111-297  TestIntervalRolling_iterate
# input: times {12,15,16,25,25,29} (25 duplicated), interval 5
# windows: (windowIndex, start, end, firstIndex, [time rows], [value rows])
# --------------------------------------------------------------------------
ITERATE_COLS = [[12, 15, 16, 25, 25, 29], [1.2, 1.5, 1.6, 2.5, 3.5, 2.9]]
ITERATE_INTERVAL = 5
_OFFSET3 = [
    (0, 8, 13, 0, [12], [1.2]),
    (1, 13, 18, 1, [15, 16], [1.5, 1.6]),
    (2, 18, 23, 3, [], []),
    (3, 23, 28, 3, [25, 25], [2.5, 3.5]),
    (4, 28, 33, 5, [29], [2.9]),
]
ITERATE = [
    ("no option", dict(offset=0, inclusive=False), [
        (0, 10, 15, 0, [12], [1.2]),
        (1, 15, 20, 1, [15, 16], [1.5, 1.6]),
        (2, 20, 25, 3, [], []),
        (3, 25, 30, 3, [25, 25, 29], [2.5, 3.5, 2.9]),
    ]),
    ("with inclusive windows", dict(offset=0, inclusive=True), [
        (0, 10, 15, 0, [12, 15], [1.2, 1.5]),
        (1, 15, 20, 1, [15, 16], [1.5, 1.6]),
        (2, 20, 25, 3, [25], [2.5]),
        (3, 25, 30, 3, [25, 25, 29], [2.5, 3.5, 2.9]),
    ]),
    ("with offset falling before first point", dict(offset=1, inclusive=False), [
        (0, 11, 16, 0, [12, 15], [1.2, 1.5]),
        (1, 16, 21, 2, [16], [1.6]),
        (2, 21, 26, 3, [25, 25], [2.5, 3.5]),
        (3, 26, 31, 5, [29], [2.9]),
    ]),
    ("with offset falling at first point", dict(offset=2, inclusive=False), [
        (0, 12, 17, 0, [12, 15, 16], [1.2, 1.5, 1.6]),
        (1, 17, 22, 3, [], []),
        (2, 22, 27, 3, [25, 25], [2.5, 3.5]),
        (3, 27, 32, 5, [29], [2.9]),
    ]),
    ("with offset falling after first point", dict(offset=3, inclusive=False), _OFFSET3),
    ("offset > interval", dict(offset=8, inclusive=False), _OFFSET3),
    ("offset == interval", dict(offset=5, inclusive=False), [
        (0, 10, 15, 0, [12], [1.2]),
        (1, 15, 20, 1, [15, 16], [1.5, 1.6]),
        (2, 20, 25, 3, [], []),
        (3, 25, 30, 3, [25, 25, 29], [2.5, 3.5, 2.9]),
    ]),
    ("offset < 0", dict(offset=-2, inclusive=False), _OFFSET3),
]

# --------------------------------------------------------------------------
# rolling/aggregation/core_test.go:24-53 shared fixtures (row based: time, value)
# --------------------------------------------------------------------------
EMPTY_ROWS = []
NIL_ROWS = [(10, N), (11, N), (20, N)]
SPARSE_FLOAT_ROWS = [
    (10, 10.0),  # partially valid window
    (11, N),
    (20, N),     # only invalid window
    # empty window
    (40, N),     # partially valid with start of window invalid
    (41, 10.0),
    (50, 10.0),  # valid with two values on start of window
    (51, 20.0),
    (61, 10.0),  # valid with two values NOT on start of window
    (69, 20.0),
]
FIXTURES = {"empty": EMPTY_ROWS, "nil": NIL_ROWS, "sparse_float": SPARSE_FLOAT_ROWS}

# core_test.go:90-108 runTestCases:
#   IntervalRolling(bow, "time", 10, Options{}).Aggregate(WindowStart(time), X(value).SetTransformations(...))
# entries: (aggregation, fixture, transformation factor or None, output value type, expected rows, citation)
_T = [10, 20, 30, 40, 50, 60]


def _rows(vals):
    return list(zip(_T, vals))


AGGREGATIONS = [
    ("ArithmeticMean", "empty", None, "float64", [], "arithmeticmean_test.go:13-24"),
    ("ArithmeticMean", "sparse_float", None, "float64", _rows([10.0, N, N, 10.0, 15.0, 15.0]),
     "arithmeticmean_test.go:25-41"),
    ("Count", "empty", None, "int64", [], "count_test.go:13-24"),
    ("Count", "sparse_float", None, "int64", _rows([1, 0, 0, 1, 2, 2]), "count_test.go:25-41"),
    ("Sum", "empty", None, "float64", [], "sum_test.go:13-24"),
    ("Sum", "sparse_float", None, "float64", _rows([10.0, 0.0, 0.0, 10.0, 30.0, 30.0]), "sum_test.go:25-41"),
    ("Min", "empty", None, "float64", [], "minmax_test.go:13-24"),
    ("Min", "sparse_float", None, "float64", _rows([10.0, N, N, 10.0, 10.0, 10.0]), "minmax_test.go:25-41"),
    ("Max", "empty", None, "float64", [], "minmax_test.go:88-99"),
    ("Max", "sparse_float", None, "float64", _rows([10.0, N, N, 10.0, 20.0, 20.0]), "minmax_test.go:100-116"),
    ("First", "empty", None, "float64", [], "firstlast_test.go:13-24"),
    ("First", "sparse_float", None, "float64", _rows([10.0, N, N, 10.0, 10.0, 10.0]), "firstlast_test.go:25-41"),
    ("Last", "empty", None, "float64", [], "firstlast_test.go:87-98"),
    ("Last", "sparse_float", None, "float64", _rows([10.0, N, N, 10.0, 20.0, 20.0]), "firstlast_test.go:99-115"),
    ("IntegralStep", "empty", None, "float64", [], "integral_test.go:14-25"),
    # Go constants: 100*0.9 = 90, 100*0.1+200*0.9 = 190, 100*0.8+200*0.1 = 100 (exact)
    ("IntegralStep", "sparse_float", None, "float64", _rows([100.0, N, N, 90.0, 190.0, 100.0]),
     "integral_test.go:26-43"),
    # integral_test.go:82-130 uses an ad-hoc closure x*0.1 — identical to transformation.Factor(0.1)
    # (factor.go:9-11).  Go folds `factor * (100.)` at run time: 0.1*100, 0.1*90, 0.1*190, 0.1*100.
    ("IntegralStep", "empty", 0.1, "float64", [], "integral_test.go:95-106"),
    ("IntegralStep", "sparse_float", 0.1, "float64",
     _rows([0.1 * 100.0, N, N, 0.1 * 90.0, 0.1 * 190.0, 0.1 * 100.0]), "integral_test.go:107-128"),
    ("IntegralTrapezoid", "empty", None, "float64", [], "integral_test.go:133-144"),
    ("IntegralTrapezoid", "sparse_float", None, "float64", _rows([N, N, N, 90.0, 15.0, 120.0]),
     "integral_test.go:145-162"),
    ("WeightedAverageStep", "empty", None, "float64", [], "weightedmean_test.go:13-24"),
    # Go constants: 10*0.9 = 9, 10*0.1+20*0.9 = 19, 10*0.8+20*0.1 = 10 (exact)
    ("WeightedAverageStep", "sparse_float", None, "float64", _rows([10.0, N, N, 9.0, 19.0, 10.0]),
     "weightedmean_test.go:25-42"),
    ("WeightedAverageStep", "nil", None, "float64", [(10, N), (20, N)], "weightedmean_test.go:43-57"),
    ("WeightedAverageLinear", "empty", None, "float64", [], "weightedmean_test.go:102-113"),
    # Go constants: 10*0.9 = 9, 15*0.1 = 1.5, 15*0.8 = 12 (exact)
    ("WeightedAverageLinear", "sparse_float", None, "float64", _rows([N, N, N, 9.0, 1.5, 12.0]),
     "weightedmean_test.go:114-131"),
]

# --------------------------------------------------------------------------
# rolling/aggregation_test.go:12-123  TestIntervalRolling_Aggregate (driver)
# input times {10,15,16,25,29}, values {1.0,1.5,1.6,2.5,2.9}, interval 10.
# The reference test uses ad-hoc closures (w.FirstValue; float64(NumRows);
# 2*float64(NumRows)).  The literal oracle replays them verbatim; the GPU path
# replays the same column plumbing with WindowStart / Count (int64) instead.
# --------------------------------------------------------------------------
AGG_DRIVER_COLS = [[10, 15, 16, 25, 29], [1.0, 1.5, 1.6, 2.5, 2.9]]
AGG_DRIVER = [
    # (name, [(kind, input, rename)], expected names, expected types, expected cols)
    ("keep columns", [("time", TIME, None), ("nrows", VALUE, None)],
     [TIME, VALUE], ["int64", "float64"], [[10, 20], [3.0, 2.0]]),
    ("swap columns", [("nrows", VALUE, None), ("time", TIME, None)],
     [VALUE, TIME], ["float64", "int64"], [[3.0, 2.0], [10, 20]]),
    ("rename columns", [("time", TIME, "a"), ("nrows", VALUE, "b")],
     ["a", "b"], ["int64", "float64"], [[10, 20], [3.0, 2.0]]),
    ("less than in original", [("time", TIME, None)], [TIME], ["int64"], [[10, 20]]),
    ("more than in original", [("time", TIME, None), ("double", VALUE, "double"), ("nrows", VALUE, None)],
     [TIME, "double", VALUE], ["int64", "float64", "float64"], [[10, 20], [6.0, 4.0], [3.0, 2.0]]),
]
AGG_DRIVER_ERRORS = [
    ("missing interval colIndex", [("nrows", VALUE, None)],
     "intervalRolling.indexedAggregations: must keep interval column 'time'"),
    ("invalid colIndex", [("time", TIME, None), ("nil", "-", None)],
     "intervalRolling.indexedAggregations: no column '-'"),
]

# --------------------------------------------------------------------------
# rolling/interpolation_test.go:11-101  TestIntervalRollingIter_Interpolate (driver)
# closures: time -> w.FirstValue, value -> constant 9.9 ; input {10,13},{1.0,1.3}, interval 2
# --------------------------------------------------------------------------
INTERP_DRIVER = [
    ("empty bow", [[], []], 0, [[], []]),                                              # :49-63
    ("no options", [[10, 13], [1.0, 1.3]], 0, [[10, 12, 13], [1.0, 9.9, 1.3]]),         # :65-82
    ("with offset", [[10, 13], [1.0, 1.3]], 1, [[9, 10, 11, 13], [9.9, 1.0, 9.9, 1.3]]),  # :84-100
]
INTERP_DRIVER_ERRORS = [
    ("invalid input type", "intervalRolling.validateInterpolation: accepts types [int64 bool], got type float64"),
    ("missing interval column", "must keep interval column 'time'"),
]

# --------------------------------------------------------------------------
# rolling/interpolation/*_test.go — Interpolate(WindowStart(time), X(value)), interval 2
# (name, X, input rows, offset, expected rows, citation)
# --------------------------------------------------------------------------
_ASC = [(10, 10.0), (15, 15.0), (17, 17.0)]
_DESC = [(10, 30.0), (15, 25.0), (17, 24.0)]
_TWO = [(10, 1.0), (13, 1.3)]
INTERPOLATIONS = [
    ("linear asc no options", "Linear", _ASC, 0,
     [(10, 10.0), (12, 12.0), (14, 14.0), (15, 15.0), (16, 16.0), (17, 17.0)], "linear_test.go:26-45"),
    ("linear asc with offset", "Linear", _ASC, 3,
     [(9, N), (10, 10.0), (11, 11.0), (13, 13.0), (15, 15.0), (17, 17.0)], "linear_test.go:47-66"),
    ("linear desc no options", "Linear", _DESC, 0,
     [(10, 30.0), (12, 28.0), (14, 26.0), (15, 25.0), (16, 24.5), (17, 24.0)], "linear_test.go:78-97"),
    ("linear desc with offset", "Linear", _DESC, 3,
     [(9, N), (10, 30.0), (11, 29.0), (13, 27.0), (15, 25.0), (17, 24.0)], "linear_test.go:99-118"),
    ("stepprevious no options", "StepPrevious", _TWO, 0,
     [(10, 1.0), (12, 1.0), (13, 1.3)], "stepprevious_test.go:19-44"),
    ("stepprevious with offset", "StepPrevious", _TWO, 1,
     [(9, N), (10, 1.0), (11, 1.0), (13, 1.3)], "stepprevious_test.go:106-132"),
    ("stepprevious with nils", "StepPrevious", [(10, 1.0), (11, N), (13, N), (15, 1.5)], 0,
     [(10, 1.0), (11, N), (12, 1.0), (13, N), (14, 1.0), (15, 1.5)], "stepprevious_test.go:134-166"),
    ("none no options", "None", _TWO, 0, [(10, 1.0), (12, N), (13, 1.3)], "none_test.go:24-42"),
    ("none with offset", "None", _TWO, 1, [(9, N), (10, 1.0), (11, N), (13, 1.3)], "none_test.go:44-63"),
]
# NOT from the reference: interpolation.StepNext does not exist upstream (the north-star names it).  Tables derived
# BY HAND from the StepPrevious tables above under the definition "value of the first row at or after the window's
# first row where time and value are valid" (Bow.GetNextValues, bowgetters.go:111-123) — they pin our own three
# implementations (literal, C, CUDA) on each other, not on the reference.
INTERPOLATIONS_STEPNEXT = [
    ("stepnext no options", "StepNext", _TWO, 0, [(10, 1.0), (12, 1.3), (13, 1.3)], "hand-derived"),
    ("stepnext with offset", "StepNext", _TWO, 1, [(9, 1.0), (10, 1.0), (11, 1.3), (13, 1.3)], "hand-derived"),
    ("stepnext with nils", "StepNext", [(10, 1.0), (11, N), (13, N), (15, 1.5)], 0,
     [(10, 1.0), (11, N), (12, 1.5), (13, N), (14, 1.5), (15, 1.5)], "hand-derived"),
    ("stepnext nothing after", "StepNext", [(10, 1.0), (13, N)], 0, [(10, 1.0), (12, N), (13, N)], "hand-derived"),
]
# StepNext ANCHORED ON THE REFERENCE: the value StepNext gives a synthetic window-start row is, by definition, what
# Bow.FillNext (bowfill.go:154-158) leaves at the window's first row — "the next valid value at or after it".  Over the
# time column t = 0, 10, .. 50 with interval 10 and offset 5 (first window start -5) window k holds exactly row k and no
# row sits on a window start, so EVERY window gets a synthetic row and its StepNext values must be row k of the
# reference's own FillNext golden table (bowfill_test.go:93-112 Int64, :269-288 Float64: FILL_CASES "Next all columns",
# whose input is FILL_ROWS; column `a` included - it is filled like the others).
STEPNEXT_VS_FILLNEXT = dict(times=[0, 10, 20, 30, 40, 50], interval=10, offset=5, rows="FILL_ROWS",
                            expected_case="Next all columns", cite="bowfill_test.go:93-112,269-288")
# rolling/interpolation/windowstart_test.go:13-64 — single-column bow {10,13}
INTERP_WINDOWSTART = [
    ("windowstart no options", [10, 13], 0, [10, 12, 13]),
    ("windowstart with offset", [10, 13], 1, [9, 10, 11, 13]),
]
INTERP_TYPE_ERRORS = [
    # linear_test.go:128-163
    ("utf8", "intervalRolling.validateInterpolation: accepts types [int64 float64], got type utf8"),
    ("bool", "intervalRolling.validateInterpolation: accepts types [int64 float64], got type bool"),
]

# --------------------------------------------------------------------------
# rolling/transformation/factor_test.go:9-35  Factor(0.1)
# --------------------------------------------------------------------------
FACTOR = [
    ("preserve nil", N, N),
    ("preserve int64", 11, 1),
    ("preserve float64", 11.0, 1.1),
]

# rolling/aggregation_test.go:125-171 TestWindow_UnsetInclusive
UNSET_INCLUSIVE = dict(cols=[[1, 2], [1, 2]], first_value=0, last_value=2, expected_cols=[[1], [1]])


# ---- aggregation.Aggregate over the whole Bow: rolling/aggregation/whole_test.go:11-290 (Int64 / Float64 cases) ----
# entries: (name, rows [(time, value)], aggregations [(constructor name, column name, output name or None)],
#           expected {"names": [...], "types": [...], "cols": [[...], ...]} or an error string, citation)
WHOLE_ROWS = [(10, 1.0), (20, 2.0), (30, 3.0)]
WHOLE_CASES = [
    ("empty bow", [], [("WindowStart", "time", None), ("ArithmeticMean", "value", None)],
     {"names": ["time", "value"], "types": ["int64", "float64"], "cols": [[], []]}, "whole_test.go:12-29"),
    ("keep columns", WHOLE_ROWS, [("WindowStart", "time", None), ("ArithmeticMean", "value", None)],
     {"names": ["time", "value"], "types": ["int64", "float64"], "cols": [[10], [2.0]]}, "whole_test.go:31-54"),
    ("swap columns", WHOLE_ROWS, [("ArithmeticMean", "value", None), ("WindowStart", "time", None)],
     {"names": ["value", "time"], "types": ["float64", "int64"], "cols": [[2.0], [10]]}, "whole_test.go:56-79"),
    ("rename columns", WHOLE_ROWS, [("WindowStart", "time", "a"), ("ArithmeticMean", "value", "b")],
     {"names": ["a", "b"], "types": ["int64", "float64"], "cols": [[10], [2.0]]}, "whole_test.go:81-104"),
    ("less columns than original", WHOLE_ROWS, [("ArithmeticMean", "value", None)],
     {"names": ["value"], "types": ["float64"], "cols": [[2.0]]}, "whole_test.go:106-127"),
    ("more columns than original", WHOLE_ROWS,
     [("ArithmeticMean", "value", "a"), ("ArithmeticMean", "value", "b"), ("ArithmeticMean", "value", "c")],
     {"names": ["a", "b", "c"], "types": ["float64"] * 3, "cols": [[2.0], [2.0], [2.0]]}, "whole_test.go:129-154"),
    ("invalid column", WHOLE_ROWS, [("WindowStart", "-", None)], "column aggregation 0: no column '-'",
     "whole_test.go:156-168"),
    ("float", WHOLE_ROWS, [("WindowStart", "time", None), ("ArithmeticMean", "value", None)],
     {"names": ["time", "value"], "types": ["int64", "float64"], "cols": [[10], [2.0]]}, "whole_test.go:170-193"),
    ("float only nil", [(10, N), (20, N), (30, N)],
     [("WindowStart", "time", None), ("WeightedAverageLinear", "value", None)],
     {"names": ["time", "value"], "types": ["int64", "float64"], "cols": [[10], [N]]}, "whole_test.go:194-217"),
]


# ---- whole-column fills: bowfill_test.go:11-380 (newFreshBow with Int64 and with Float64 columns a..e) -----------
FILL_ROWS = [
    [20, 6, 30, 400, -10],
    [13, N, N, N, N],
    [10, 4, 10, 10, -5],
    [0, N, 3, 4, 0],
    [N, N, N, N, N],
    [-2, 1, N, N, -8],
]
# entries: (name, method, args, expected rows for Int64 columns, expected rows for Float64 columns, citation)
FILL_CASES = [
    ("Mean one column", "FillMean", (1,),
     [[20, 6, 30, 400, -10], [13, 5, N, N, N], [10, 4, 10, 10, -5], [0, 3, 3, 4, 0], [N, 3, N, N, N], [-2, 1, N, N, -8]],
     [[20, 6, 30, 400, -10], [13, 5, N, N, N], [10, 4, 10, 10, -5], [0, 2.5, 3, 4, 0], [N, 2.5, N, N, N],
      [-2, 1, N, N, -8]], "bowfill_test.go:30-49,206-225"),
    ("Mean all columns", "FillMean", (),
     [[20, 6, 30, 400, -10], [13, 5, 20, 205, -8], [10, 4, 10, 10, -5], [0, 3, 3, 4, 0], [-1, 3, N, N, -4],
      [-2, 1, N, N, -8]],
     [[20, 6, 30, 400, -10], [13, 5, 20, 205, -7.5], [10, 4, 10, 10, -5], [0, 2.5, 3, 4, 0], [-1, 2.5, N, N, -4],
      [-2, 1, N, N, -8]], "bowfill_test.go:51-70,227-246"),
    ("Next one column", "FillNext", (1,),
     [[20, 6, 30, 400, -10], [13, 4, N, N, N], [10, 4, 10, 10, -5], [0, 1, 3, 4, 0], [N, 1, N, N, N], [-2, 1, N, N, -8]],
     None, "bowfill_test.go:72-91,248-267"),
    ("Next all columns", "FillNext", (),
     [[20, 6, 30, 400, -10], [13, 4, 10, 10, -5], [10, 4, 10, 10, -5], [0, 1, 3, 4, 0], [-2, 1, N, N, -8],
      [-2, 1, N, N, -8]], None, "bowfill_test.go:93-112,269-288"),
    ("Previous one column", "FillPrevious", (1,),
     [[20, 6, 30, 400, -10], [13, 6, N, N, N], [10, 4, 10, 10, -5], [0, 4, 3, 4, 0], [N, 4, N, N, N], [-2, 1, N, N, -8]],
     None, "bowfill_test.go:114-133,290-309"),
    ("Previous all columns", "FillPrevious", (),
     [[20, 6, 30, 400, -10], [13, 6, 30, 400, -10], [10, 4, 10, 10, -5], [0, 4, 3, 4, 0], [0, 4, 3, 4, 0],
      [-2, 1, 3, 4, -8]], None, "bowfill_test.go:135-154,311-330"),
    ("Linear refCol a toFillCol b (desc)", "FillLinear", (0, 1),
     [[20, 6, 30, 400, -10], [13, 5, N, N, N], [10, 4, 10, 10, -5], [0, 2, 3, 4, 0], [N, N, N, N, N], [-2, 1, N, N, -8]],
     [[20, 6, 30, 400, -10], [13, 4.6, N, N, N], [10, 4, 10, 10, -5], [0, 1.5, 3, 4, 0], [N, N, N, N, N],
      [-2, 1, N, N, -8]], "bowfill_test.go:156-174,332-351"),
    ("Linear refCol a toFillCol e (asc)", "FillLinear", (0, 4),
     [[20, 6, 30, 400, -10], [13, N, N, N, -7], [10, 4, 10, 10, -5], [0, N, 3, 4, 0], [N, N, N, N, N], [-2, 1, N, N, -8]],
     [[20, 6, 30, 400, -10], [13, N, N, N, -6.5], [10, 4, 10, 10, -5], [0, N, 3, 4, 0], [N, N, N, N, N],
      [-2, 1, N, N, -8]], "bowfill_test.go:176-195,353-372"),
    ("Linear refCol not sorted", "FillLinear", (4, 1), "error", "error", "bowfill_test.go:197-202,374-379"),
]


# ---- Bow.DropNils: bow_test.go:165-282 ; Bow.IsColSorted: bowassertion_test.go:11-55 ------------------------------
DROP_HOLED = [[N, 200, 300, 400], [110, N, 330, 440], [111, N, 333, N]]     # columns a, b, c
DROP_CASES = [
    ("empty bow", [[]], (), [[]], "bow_test.go:183-199"),
    ("unchanged without nil", [[100, 200, 300, 400], [110, 220, 330, 440], [111, 222, 333, 444]], (),
     [[100, 200, 300, 400], [110, 220, 330, 440], [111, 222, 333, 444]], "bow_test.go:201-206"),
    ("drop on all columns", DROP_HOLED, (), [[300], [330], [333]], "bow_test.go:217-231"),
    ("drop listed = default", DROP_HOLED, (1, 2, 0), [[300], [330], [333]], "bow_test.go:208-215"),
    ("drop on one column", DROP_HOLED, (1,), [[N, 300, 400], [110, 330, 440], [111, 333, N]], "bow_test.go:233-247"),
    ("drop consecutively at start/middle/end", [[N, N, 1, N, N, 2, N, N]], (), [[1, 2]], "bow_test.go:249-268"),
]
SORTED_ROWS = [[-2, 1, N, N, -8], [0, N, 3, 4, 0], [1, N, N, 120, N], [10, 4, 10, 10, -5], [13, N, N, N, N],
               [20, 6, 30, 400, -10]]
SORTED_EXPECTED = [True, True, True, False, False]


# ---- Bow.SortByCol: bowsort_test.go:11-208 (Boolean / String columns of the upstream tables left out) ----------------
# (name, column types, input rows, sort column, expected rows | "same" | "error", cite)
SORT_CASES = [
    ("sorted", "iff", [[10, 2.4, 3.1], [11, 2.8, 5.9], [12, 2.9, 7.5], [13, 3.9, 13.4]], 0, "same", "bowsort_test.go:12-27"),
    ("unsorted with all types", "iif", [[10, 2, 3.1], [11, 2, 5.9], [13, 3, 13.4], [12, 2, 7.5]], 0,
     [[10, 2, 3.1], [11, 2, 5.9], [12, 2, 7.5], [13, 3, 13.4]], "bowsort_test.go:29-53"),
    ("unsorted with different cols", "ffi", [[2.4, 3.1, 10], [2.8, 5.9, 11], [3.9, 13.4, 13], [2.9, 7.5, 12]], 2,
     [[2.4, 3.1, 10], [2.8, 5.9, 11], [2.9, 7.5, 12], [3.9, 13.4, 13]], "bowsort_test.go:55-79"),
    ("unsorted with nil values", "iif", [[10, 5, N], [11, 2, 56.], [13, N, 13.4], [12, -1, N]], 0,
     [[10, 5, N], [11, 2, 56.], [12, -1, N], [13, N, 13.4]], "bowsort_test.go:81-105"),
    ("sorted in desc order", "iff", [[13, 3.9, 13.4], [12, 2.9, 7.5], [11, 2.8, 5.9], [10, 2.4, 3.1]], 0,
     [[10, 2.4, 3.1], [11, 2.8, 5.9], [12, 2.9, 7.5], [13, 3.9, 13.4]], "bowsort_test.go:107-131"),
    ("duplicate values in sort by column", "iff", [[13, 3.9, 13.4], [12, 2.9, 7.5], [12, 2.8, 5.9], [10, 2.4, 3.1]], 0,
     [[10, 2.4, 3.1], [12, 2.9, 7.5], [12, 2.8, 5.9], [13, 3.9, 13.4]], "bowsort_test.go:133-157"),
    ("empty bow", "if", [], 0, "same", "bowsort_test.go:159-169"),
    ("with metadata", "if", [[1, .1], [3, .3], [2, .2]], 0, [[1, .1], [2, .2], [3, .3]], "bowsort_test.go:171-188"),
    ("ERR: nil values in sort by column", "iff", [[13, 3.9, 13.4], [12, 2.9, 7.5], [N, 2.8, 5.9], [10, 2.4, 3.1]], 0,
     "error", "bowsort_test.go:190-204"),
]
