from tensorflow import (reshape, concat, stack, unstack, slice, tile, transpose, gather, where, zeros, ones,  # noqa: F401,A004
                        squeeze, expand_dims, identity)
