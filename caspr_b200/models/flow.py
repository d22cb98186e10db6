"""Point-CNF construction helpers (mirror of reference caspr/models/flow.py)."""
from .cnf import CNF, SequentialFlow, ODEfunc, ODEnet, MovingBatchNorm1d
from .latent_ode_model import LatentODE


def count_nfe(model):
    """flow.py:10-22: sum of num_evals() over CNF and LatentODE sub-modules."""
    total = [0]

    def acc(module):
        if isinstance(module, (CNF, LatentODE)):
            total[0] += module.num_evals()

    model.apply(acc)
    return total[0]


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def build_model(args, input_dim, hidden_dims, context_dim, num_blocks, conditional):
    """flow.py:44-76."""
    def build_cnf():
        diffeq = ODEnet(hidden_dims=hidden_dims, input_shape=(input_dim,), context_dim=context_dim,
                        layer_type=args.layer_type, nonlinearity=args.nonlinearity)
        return CNF(odefunc=ODEfunc(diffeq=diffeq), T=args.time_length, train_T=args.train_T,
                   conditional=conditional, solver=args.solver, use_adjoint=args.use_adjoint,
                   atol=args.atol, rtol=args.rtol)

    chain = [build_cnf() for _ in range(num_blocks)]
    if args.batch_norm:
        chain = [MovingBatchNorm1d(input_dim)] + chain + [MovingBatchNorm1d(input_dim)]
    return SequentialFlow(chain, use_bn=args.batch_norm)


def get_point_cnf(args):
    """flow.py:79-83.  The reference hard-calls ``.cuda()`` here; we do the same whenever a GPU is
    visible (on a GPU-less machine the module can still be built to inspect its state_dict)."""
    import torch
    dims = tuple(map(int, args.dims.split('-')))
    model = build_model(args, args.input_dim, dims, args.zdim, args.num_blocks, True)
    return model.cuda() if torch.cuda.is_available() else model


class PointCNFArgs():
    """flow.py:86-100."""

    def __init__(self):
        self.input_dim = 3
        self.dims = '512-512-512'
        self.zdim = 512
        self.num_blocks = 1
        self.layer_type = 'concatsquash'
        self.nonlinearity = 'softplus'
        self.time_length = 0.5
        self.train_T = True
        self.solver = 'dopri5'
        self.use_adjoint = True
        self.atol = 1e-5
        self.rtol = 1e-5
        self.batch_norm = True
