"""numpy (float64) evaluation of the flat plan descriptor the C-ABI executor consumes
(``include/bayesic_b200.h``: bb_node_desc).  TEST INFRASTRUCTURE: it lets the parity
tests check the CUDA executor node-for-node on exactly the descriptor it was given,
independently of the expression-level oracle in ``semantics.py``.

Node semantics restate the reference's executor vocabulary: ``_sum`` algebra.py:1290-1291,
``_mul`` :1302-1306, ``_dimshuffle`` :1318-1319, ``_tensordot`` :1161-1171 (declared
semantics, batch + X_other + Y_other), ``_diagonal`` :1405-1407, elementwise :1435-1448.
The fused nodes are defined by the sub-trees they replace (``backend/lowering.py``)."""
import numpy as np
from scipy.special import log_softmax

from .semantics import logdet_spd, tensordot_declared

KIND = {0: 'input', 1: 'scalar', 2: 'shape', 3: 'eye', 4: 'sum', 5: 'mul', 6: 'dimshuffle',
        7: 'tensordot', 8: 'diagonal', 9: 'elemwise', 20: 'logsoftmax', 21: 'syrk',
        22: 'weighted_scatter', 23: 'logdet'}
OPS = {0: 'add', 1: 'mul', 2: 'log', 3: 'exp', 4: 'pow', 5: 'abs', 6: 'lgamma'}


def evaluate_descriptor(nodes, outputs, input_arrays, dtype=np.float64):
    """``nodes``: list of dicts (kind, parents, iparams, fparam); ``input_arrays``: per slot."""
    vals = []
    for node in nodes:
        kind = KIND[node['kind']]
        par = [vals[p] for p in node['parents']]
        ip = node['iparams']
        if kind == 'input':
            value = np.asarray(input_arrays[ip[0]])
            value = value.astype(dtype) if value.dtype.kind == 'f' else value.astype(dtype)
        elif kind == 'scalar':
            value = np.asarray(node['fparam'], dtype=dtype)
        elif kind == 'shape':
            value = np.asarray(par[0].shape[ip[0]], dtype=dtype)
        elif kind == 'eye':
            value = np.eye(int(round(float(par[0]))), dtype=dtype)
        elif kind == 'sum':
            value = par[0].sum(axis=tuple(ip))
        elif kind == 'mul':
            value = par[0]
            for p in par[1:]:
                value = value * p
        elif kind == 'dimshuffle':
            value = np.transpose(par[0], [a for a in ip if a >= 0])
            value = value[tuple(np.newaxis if a < 0 else slice(None) for a in ip)]
        elif kind == 'diagonal':
            value = np.diagonal(par[0], 0, ip[0], ip[1])
        elif kind == 'tensordot':
            n_dot, n_batch = ip[0], ip[1]
            rest = ip[2:]
            x_dot, y_dot = rest[:n_dot], rest[n_dot:2 * n_dot]
            x_batch = rest[2 * n_dot:2 * n_dot + n_batch]
            y_batch = rest[2 * n_dot + n_batch:]
            value = tensordot_declared(par[0], par[1], x_dot, y_dot, x_batch, y_batch)
        elif kind == 'elemwise':
            op = OPS[ip[0]]
            if op == 'add':
                value = par[0]
                for p in par[1:]:
                    value = value + p
            elif op == 'mul':
                value = par[0]
                for p in par[1:]:
                    value = value * p
            elif op == 'log':
                value = np.log(par[0])
            elif op == 'exp':
                value = np.exp(par[0])
            elif op == 'pow':
                value = np.power(par[0], par[1])
            elif op == 'lgamma':
                from scipy.special import gammaln
                value = gammaln(par[0])
            else:
                value = np.abs(par[0])
        elif kind == 'logsoftmax':
            value = log_softmax(par[0], axis=-1)
        elif kind == 'syrk':
            value = par[0].T @ par[0]
        elif kind == 'weighted_scatter':
            value = np.einsum('nk,nd,ne->kde', par[0], par[1], par[1])
        elif kind == 'logdet':
            value = logdet_spd(par[0])
        else:
            raise ValueError(kind)
        vals.append(np.asarray(value))
    return [vals[o] for o in outputs]
