from tensorflow import constant_initializer, random_uniform_initializer  # noqa: F401
