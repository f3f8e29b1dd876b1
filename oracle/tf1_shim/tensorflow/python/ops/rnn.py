from tensorflow.contrib.rnn import static_rnn  # noqa: F401
