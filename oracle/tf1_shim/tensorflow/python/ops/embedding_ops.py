from tensorflow.nn import embedding_lookup  # noqa: F401
