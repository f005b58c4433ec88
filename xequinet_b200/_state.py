"""Which gradients does THIS backward pass actually want?

`ctx.needs_input_grad` of a custom autograd formula is fixed when the forward runs, so it cannot tell that
`torch.autograd.grad(E, [pos], create_graph=True)` -- the force pass of nn/basic.py:150-156 -- asks for d/dpos only:
without more information every weight-gradient GEMM, column sum and rbf weight-gradient kernel of the model would run
(and be recorded for double differentiation) only to be thrown away.  The autograd engine knows which nodes of the
current graph task it is going to execute; `input_wanted(ctx, i)` asks it whether the producer of input i is one of
them.  Stateless (nothing global, nothing thread-local), exact per backward call, and valid on the engine's device
threads, where Python thread-local state of the caller is not visible."""
from __future__ import annotations

import torch


def input_wanted(ctx, i: int) -> bool:
    """True when the gradient with respect to input i of the autograd.Function node `ctx` is needed by the graph
    task that is running this backward formula."""
    if not ctx.needs_input_grad[i]:
        return False
    fn = ctx.next_functions[i][0]
    if fn is None:
        return False
    try:
        return bool(torch._C._will_engine_execute_node(fn))
    except RuntimeError:
        # a leaf that autograd.grad() was asked to differentiate with respect to (the engine refuses the query for
        # those): it is wanted by construction
        return True
