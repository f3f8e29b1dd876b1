from tensorflow.nn import sparse_softmax_cross_entropy_with_logits, softmax, log_softmax, relu, dropout, top_k  # noqa: F401
