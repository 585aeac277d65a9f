"""dgl.backend.asnumpy (utils.py:110, the unused 'RW' positional-encoding branch)."""


def asnumpy(t):
    return t.detach().cpu().numpy()
