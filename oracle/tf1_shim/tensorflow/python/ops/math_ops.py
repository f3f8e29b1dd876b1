from tensorflow import (add_n, reduce_sum, reduce_mean, reduce_max, cast, matmul, add, subtract, multiply, div,  # noqa: F401
                        exp, log, square, sqrt, sigmoid, tanh, maximum, minimum, range)  # noqa: A004
