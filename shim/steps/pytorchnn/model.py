"""Drop-in for the reference's ``steps/pytorchnn/model.py``: put this directory first on PYTHONPATH (the pipeline
already runs its drivers with ``PYTHONPATH=steps/pytorchnn``, lmrescore_nbest_pytorchnn_cuda.sh:200) and every
``import model`` (compute_sentence_scores_bayes_jianwei.py:373, train.py:22) resolves to the B200 classes:
same names, positional constructor orders, ``forward`` / ``init_hidden`` / ``kl_divergence`` surface and
``state_dict`` keys, computed by libbayeslm_b200.so."""
import os
import sys

_ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", ".."))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from bayeslms_b200.model import (  # noqa: E402,F401
    BayesLinear, BayesMultiheadAttention, BayesRNNModel, BayesTransformerEncoderLayer, BayesTransformerModel,
    Bayes2LSTM, GPLSTM, GPLSTMCell, GPNN, GaussRNNModel, GaussTransformerEncoderLayer, GaussTransformerModel,
    MultiheadAttention, PositionalEncoding, RNNModel, StandardTransformerEncoderLayer, TransformerModel, VLSTMCell,
    VNN, VTransformerEncoderLayer, VTransformerModel, VariationalLSTM, VariationalRNNModel)
