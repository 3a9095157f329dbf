"""Shared helpers for the tests."""
import torch


def build_from_cfg(cfg, device=None):
    from bayeslms_b200 import model as M
    fam = cfg["family"]
    if fam == "bayes_tm":
        net = M.BayesTransformerModel(cfg["ntoken"], cfg["ninp"], cfg["nhead"], cfg["nhid"], cfg["nlayers"], 0.5, True,
                                      cfg["bayes_pos"])
    elif fam == "gauss_tm":
        net = M.GaussTransformerModel(cfg["ntoken"], cfg["ninp"], cfg["nhead"], cfg["nhid"], cfg["nlayers"], 0.5, True,
                                      cfg["gauss_pos"])
    elif fam == "v_tm":
        net = M.VTransformerModel(cfg["ntoken"], cfg["ninp"], cfg["nhead"], cfg["nhid"], cfg["nlayers"], 0.5, True,
                                  cfg["v_pos"])
    elif fam == "bayes_lstm":
        net = M.BayesRNNModel("LSTM", cfg["ntoken"], cfg["ninp"], cfg["nhid"], cfg["nlayers"], 0.5, True,
                              cfg["bayes_pos"])
    elif fam == "gauss_lstm":
        net = M.GaussRNNModel("LSTM", cfg["ntoken"], cfg["ninp"], cfg["nhid"], cfg["nlayers"], 0.5, False, cfg["gauss_pos"])
    elif fam == "v_lstm":
        net = M.VariationalRNNModel("LSTM", cfg["ntoken"], cfg["ninp"], cfg["nhid"], cfg["nlayers"], 0.5, True, cfg["v_pos"])
    elif fam == "std_tm":
        net = M.TransformerModel(cfg["ntoken"], cfg["ninp"], cfg["nhead"], cfg["nhid"], cfg["nlayers"], 0.5,
                                 cfg.get("activation", "gelu"), True)
    elif fam == "std_lstm":
        net = M.RNNModel("LSTM", cfg["ntoken"], cfg["ninp"], cfg["nhid"], cfg["nlayers"], 0.5, True)
    else:
        raise ValueError(fam)
    return net if device is None else net.to(device)


def load_golden_model(rec, device):
    net = build_from_cfg(rec["cfg"])
    missing, unexpected = net.load_state_dict(rec["state_dict"], strict=False)
    assert not unexpected and all(k == "pos_encoder.pe" for k in missing), (missing, unexpected)
    return net.to(device).eval()
