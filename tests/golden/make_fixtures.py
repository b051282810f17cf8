"""Writes tests/golden/reference_fixtures.json.

The reference is Julia and cannot run in this image, so there is nothing to "execute" to generate
vectors.  What the reference's own test-suite pins are LITERAL inputs and expected outputs; they
are transcribed here (with file:line) and the expected values the Julia tests compute with plain
integer arithmetic (`mod.(A*B, N)` etc.) are evaluated below with exact python ints -- NOT with
oracle/ -- so that the oracle can then be checked against them.

Run:  python tests/golden/make_fixtures.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def matmul(A, B):
    return [[sum(A[i][k] * B[k][j] for k in range(len(B))) for j in range(len(B[0]))] for i in range(len(A))]


def fmod(M, N):  # Julia mod (floored)
    return [[x % N for x in r] for r in M]


def ew(f, A, B):
    return [[f(a, b) for a, b in zip(ra, rb)] for ra, rb in zip(A, B)]


fx = {}

# test/CuModMatrix/de_rham_test.jl:3-37 -- is_invertible_with_inverse(A mod 7) is true, A*B == I
fx["de_rham"] = {
    "cite": "test/CuModMatrix/de_rham_test.jl:3-37",
    "N": 7,
    "A": [
        [0, 0, 0, 0, 0, 0, 0, -3, 0, 0],
        [0, 0, 0, 0, 0, 0, 0, 0, -3, 0],
        [0, 1, 0, 0, 0, 0, 0, 0, 0, 0],
        [0, 0, 1, 0, 0, 0, 0, 0, 0, 0],
        [0, -2, 0, 0, 0, 0, 0, 0, 0, -3],
        [0, 0, -2, 0, 2, 0, 0, 0, 0, 0],
        [0, 0, 0, 1, 0, 2, 0, 0, 0, 0],
        [0, -3, 0, -2, 0, 0, 0, -1, 0, 0],
        [0, 0, -3, 0, 0, 0, 2, 0, -1, 0],
        [1, 0, 0, -3, 0, 0, 0, 0, 0, -1],
    ],
    "invertible": True,
    "A_times_inverse": [[1 if i == j else 0 for j in range(10)] for i in range(10)],
}

# test/CuModMatrix/matmul_operations_test.jl:14-67
A = [[1, 2, 3], [4, 5, 6]]
B = [[7, 8], [9, 10], [0, 1]]
fx["matmul_2x3_3x2"] = {
    "cite": "test/CuModMatrix/matmul_operations_test.jl:14-67",
    "A": A, "B": B, "N": 11,
    "C_literal": [[58, 64], [139, 154]],          # :17
    "C_mod11": fmod(matmul(A, B), 11),            # == [3 9; 7 0] per the comment at :17
    "override_N": 7, "C_mod7": fmod(matmul(A, B), 7),  # :55-65
}
assert fx["matmul_2x3_3x2"]["C_mod11"] == [[3, 9], [7, 0]]
# the literal at :17 is A*[7 8;9 10;11 12]; it equals A*B only mod 11 (11=0, 12=1), which is all the test uses
assert fmod([[58, 64], [139, 154]], 11) == fmod(matmul(A, B), 11)

# test/CuModMatrix/matmul_operations_test.jl:79-126
B2 = [[7, 8], [9, 10], [11, 12]]
fx["matmul_inplace"] = {
    "cite": "test/CuModMatrix/matmul_operations_test.jl:79-126",
    "A": A, "B": B2, "N": 9, "C_mod9": fmod(matmul(A, B2), 9),
    "override_N": 3, "C_mod3": fmod(matmul(A, B2), 3),
}

# test/CuModMatrix/basic_operations_test.jl:18-117
A3 = [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
B3 = [[9, 8, 7], [6, 5, 4], [3, 2, 1]]
N = 11
fx["basic_3x3"] = {
    "cite": "test/CuModMatrix/basic_operations_test.jl:18-117",
    "A": A3, "B": B3, "N": N, "scalar": 3,
    "add": fmod(ew(lambda a, b: a + b, A3, B3), N),
    "sub": fmod(ew(lambda a, b: a - b, A3, B3), N),
    "matmul": fmod(matmul(A3, B3), N),
    "elementwise_multiply": fmod(ew(lambda a, b: a * b, A3, B3), N),
    "scalar_add": fmod([[3 + a for a in r] for r in A3], N),
    "scalar_sub": fmod([[a - 3 for a in r] for r in A3], N),
    "scalar_mul": fmod([[3 * a for a in r] for r in A3], N),
    "negate": fmod([[-a for a in r] for r in A3], N),
    # :189-200  G^2 and G^0
    "pow2": fmod(matmul(A3, A3), N),
    "pow0": [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
}
assert fx["basic_3x3"]["add"] == [[10] * 3] * 3

# test/CuModMatrix/permutation_test.jl:3-55
fx["permutation_3x3"] = {
    "cite": "test/CuModMatrix/permutation_test.jl:3-55",
    "A": A3, "N": 11, "P": [[2, 3]],
    "col_perm": [[1, 3, 2], [4, 6, 5], [7, 9, 8]],
    "row_perm": [[1, 2, 3], [7, 8, 9], [4, 5, 6]],
}

# test/CuModMatrix/triangular_test.jl:21-26, :72-77
fx["triangular_2x2"] = {
    "cite": "test/CuModMatrix/triangular_test.jl:21-26,72-77",
    "N": 7, "upper": [[1, 3], [0, 2]], "lower": [[1, 0], [3, 2]],
}

# test/CuModMatrix/inplace_operations_test.jl:216-255: fill!(F,122.0) -> 1 mod 11
fx["fill_122_mod_11"] = {"cite": "test/CuModMatrix/inplace_operations_test.jl:216-255", "N": 11, "value": 122, "expect": 122 % 11}

# test/KaratsubaMatrix/basic_operations_test.jl:75-133 -- property with N1=13^4, N2=13^3, n=500
fx["karatsuba_params"] = {"cite": "test/KaratsubaMatrix/basic_operations_test.jl:75-133", "N1": 13 ** 4, "N2": 13 ** 3, "n": 500}

# test/CuModMatrix/stripe_mul_test.jl:3-50 -- property cases
fx["stripe_cases"] = {
    "cite": "test/CuModMatrix/stripe_mul_test.jl:3-50",
    "cases": [
        {"n": 100, "N": 2 ** 11, "lo": 1, "hi": 100},
        {"n": 100, "N": 11 ** 3, "lo": 1, "hi": 100, "poke": [[5, 5, 11 ** 3 - 5, "A"], [95, 95, 11 ** 3 - 23, "B"]]},
        {"n": 3003, "N": 11 ** 3, "lo": 1, "hi": 11 ** 3, "poke": [[5, 5, 11 ** 3 - 5, "A"], [95, 95, 11 ** 3 - 23, "B"]]},
    ],
}

with open(os.path.join(HERE, "reference_fixtures.json"), "w") as f:
    json.dump(fx, f, indent=1, sort_keys=True)
print("wrote", os.path.join(HERE, "reference_fixtures.json"))
