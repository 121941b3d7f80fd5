"""Unsupervised EA trainer: the schedule of the reference's run/train_unsup_ea.py:43-159 (Wasserstein phase,
then refinement on mutual nearest neighbours with hard negatives) over the eagraft models.  The timed
benchmark step in bench.py is the body of the first loop below."""
import time

import numpy as np
import torch

from ..models.models_ea import UEAModel
from ..utils.data_utils import load_seperate_data_ea
from ..utils.eval_utils import eval_at_1, format_metrics


def train_unsup_ea(args, data=None, log=print):
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    args.device = 'cuda:' + str(args.cuda)
    data = data if data is not None else load_seperate_data_ea(args, args.data_root)
    args.n_nodes, args.feat_dim = data['x'].shape
    args.n_classes = args.feat_dim
    args.data = data
    model = UEAModel(args).to(args.device)
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=int(args.lr_reduce_freq), gamma=float(args.gamma))
    x, adj = data['x'], data['adj']
    history = {"glove_hits1": float(eval_at_1(x, data)), "wasserstein": [], "refine": []}
    log("Only Glove Hits@1: {:.4f}".format(history["glove_hits1"]))
    t_total = time.time()

    def forward():
        return model.decode(model.encode(x, adj), adj)

    for epoch in range(args.epochs):                      # minimise the (as-shipped) Wasserstein objective
        model.train()
        optimizer.zero_grad()
        outputs = forward()
        for _ in range(args.iters):
            loss = model.get_loss_wassertein(outputs, data, args.batch_size)
            loss.backward()
            optimizer.step()
            model.join_pending_solve()
        scheduler.step()
        if (epoch + 1) % args.eval_freq == 0:
            model.eval()
            with torch.no_grad():
                hits1 = float(eval_at_1(forward(), data))
            history["wasserstein"].append((float(loss), hits1))
            log('Epoch: {:04d} Loss: {:.4f} Test Hits@1: {:.4f}%'.format(epoch + 1, float(loss), hits1))

    for epoch in range(args.refine_epochs):               # refinement on pseudo-labels
        model.train()
        optimizer.zero_grad()
        outputs = forward()
        if epoch % 10 == 0:
            model.generate_pairs(outputs, data, args.batch_size)
            model.generate_neg(outputs, args.neg_num)
        loss = model.get_loss(outputs)
        loss.backward()
        optimizer.step()
        scheduler.step()
        model.eval()
        with torch.no_grad():
            hits1 = float(eval_at_1(forward(), data))
        history["refine"].append((float(loss), hits1))
        log('Refine-Epoch: {:04d} Loss: {:.4f} Test Hits@1: {:.4f}%'.format(epoch + 1, float(loss), hits1))

    model.eval()
    with torch.no_grad():
        outputs = forward()
        history["test"] = model.compute_metrics(outputs, data, 'test')
    log('Total time elapsed: {:.4f}s'.format(time.time() - t_total))
    log('Test set results: ' + format_metrics(history["test"], 'test'))
    return model, history
