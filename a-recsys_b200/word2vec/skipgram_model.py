"""Skip-gram "linear sequence" model on B200 — same constructor and step() protocol as the reference
(word2vec/skipgram_model.py:13-140, word2vec/linear_seq.py:70-120).

  training : h = dropout( mean( user_emb, mean_f item_emb_f(input_0) ) )  -> the (input, output) pair of get_next_sg
             (skipgram_model.py:86-88: only the first input placeholder feeds the training tower)
  test     : h = mean( user_emb, mean_k mean_f item_emb_f(input_k) )       (:90-99, the CBOW test tower)
Scoring, losses (ce / warp / bbpr / mw) and the optimizer are the CBOW model's: same kernels, one flag.
"""
from . import cbow_model


class Model(cbow_model.Model):
    train_first_input_only = True
