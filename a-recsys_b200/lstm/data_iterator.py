"""Batch streams for the LSTM runner (same interface as the reference's lstm/data_iterator.py:6-42:
`DataIterator(model, data_set, n_bucket, batch_size, train_buckets_scale)` with the generators
`next_random()` and `next_sequence(stop, recommend)`), built on the model's own batch builders."""
import numpy as np


class DataIterator(object):
    def __init__(self, model, data_set, n_bucket, batch_size, train_buckets_scale):
        self.model, self.data_set = model, data_set
        self.n_bucket, self.batch_size = n_bucket, batch_size
        self.train_buckets_scale = train_buckets_scale
        # cumulative share of the training sequences per bucket (only the random stream needs it)
        self._cdf = None if train_buckets_scale is None else np.asarray(train_buckets_scale, dtype=np.float64)

    def _draw_bucket(self):
        """First bucket whose cumulative share exceeds one uniform draw: buckets are visited in proportion
        to their size (lstm/data_iterator.py:14-18)."""
        return int(np.searchsorted(self._cdf, np.random.random_sample(), side='right'))

    def next_random(self):
        """Endless training stream: a size-weighted random bucket, then a batch drawn with replacement."""
        while True:
            b = self._draw_bucket()
            users, inputs, outputs, weights, _ = self.model.get_batch(self.data_set, b)
            yield users, inputs, outputs, weights, b

    def next_sequence(self, stop=False, recommend=False):
        """Deterministic sweep: every bucket in order, consecutive windows of batch_size sequences until the
        batch builder reports the bucket exhausted; one pass when `stop`, else round and round.  Deliberate difference
        from lstm/data_iterator.py:22-42: an EMPTY bucket is skipped, where the reference issues one all-padding batch
        (zero weights: it adds nothing to a loss or a recommendation, only a kernel launch); pinned by
        tests/test_host_logic_vs_reference.py::test_lstm_data_iterator_streams_match_reference."""
        fetch = self.model.get_batch_recommend if recommend else self.model.get_batch
        while True:
            for b in range(self.n_bucket):
                if len(self.data_set[b]) == 0:
                    continue
                offset, exhausted = 0, False
                while not exhausted:
                    users, inputs, outputs, weights, exhausted = fetch(self.data_set, b, start_id=offset)
                    yield users, inputs, outputs, weights, b
                    offset += self.batch_size
            if stop:
                return
