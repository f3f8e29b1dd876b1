"""Batch iterators of the LSTM runner (reference: lstm/data_iterator.py:6-42)."""
import numpy as np


class DataIterator(object):
    def __init__(self, model, data_set, n_bucket, batch_size, train_buckets_scale):
        self.data_set = data_set
        self.n_bucket = n_bucket
        self.batch_size = batch_size
        self.train_buckets_scale = train_buckets_scale
        self.model = model

    def next_random(self):
        """bucket chosen with probability proportional to its size, batch drawn with replacement."""
        while True:
            r = np.random.random_sample()
            bucket_id = min(i for i in range(len(self.train_buckets_scale)) if self.train_buckets_scale[i] > r)
            users, inputs, outputs, weights, _ = self.model.get_batch(self.data_set, bucket_id)
            yield users, inputs, outputs, weights, bucket_id

    def next_sequence(self, stop=False, recommend=False):
        bucket_id = 0
        while True:
            if bucket_id >= self.n_bucket:
                if stop:
                    break
                bucket_id = 0
            start_id = 0
            while True:
                fn = self.model.get_batch_recommend if recommend else self.model.get_batch
                if len(self.data_set[bucket_id]) == 0:
                    break
                users, inputs, outputs, weights, finished = fn(self.data_set, bucket_id, start_id=start_id)
                yield users, inputs, outputs, weights, bucket_id
                if finished:
                    break
                start_id += self.batch_size
            bucket_id += 1
