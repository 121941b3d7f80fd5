"""Supervised EA trainer: the schedule of the reference's run/train_ea.py:8-110 (margin loss, hard
negatives refreshed every 50 epochs, per-epoch Hits@k) over the eagraft models.  Unlike the reference —
whose first evaluation raises KeyError('Hits@10_l') because compute_metrics asks for top_k=[1] only
(models/models_ea.py:66,69) — early stopping here reads Hits@1 when Hits@10 is absent."""
import time

import numpy as np
import torch

from ..models.models_ea import EAModel
from ..utils.data_utils import load_data_ea
from ..utils.eval_utils import format_metrics


def _improved(best, new):
    key = 'Hits@10' if 'Hits@10_l' in new else 'Hits@1'
    return best.get(key + '_l', -1) < new[key + '_l'] or best.get(key + '_r', -1) < new[key + '_r']


def train_ea(args, data=None, log=print):
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    args.device = 'cuda:' + str(args.cuda)
    data = data if data is not None else load_data_ea(args, args.data_root)
    args.n_nodes, args.feat_dim = data['x'].shape
    args.n_classes = args.feat_dim
    args.data = data
    model = EAModel(args).to(args.device)
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=int(args.lr_reduce_freq), gamma=float(args.gamma))
    x, adj = data['x'], data['adj']
    best_val, best_test, counter = model.init_metric_dict(), None, 0
    t_total = time.time()
    for epoch in range(args.epochs):
        model.train()
        optimizer.zero_grad()
        outputs = model.decode(model.encode(x, adj), adj)
        if epoch % 50 == 0:
            model.neg_right = model.get_neg(data['train'][:, 0], outputs, args.neg_num)
            model.neg2_left = model.get_neg(data['train'][:, 1], outputs, args.neg_num)
        loss = model.get_loss(outputs, data, 'train')
        loss.backward()
        optimizer.step()
        scheduler.step()
        if (epoch + 1) % args.log_freq == 0:
            train_metrics = model.compute_metrics(outputs.detach(), data, 'train')
            log('Epoch: {:04d} loss: {:.4f} {}'.format(epoch + 1, float(loss), format_metrics(train_metrics, 'train')))
        if (epoch + 1) % args.eval_freq == 0:
            model.eval()
            with torch.no_grad():
                outputs = model.decode(model.encode(x, adj), adj)
                val_metrics = model.compute_metrics(outputs, data, 'val')
            if _improved(best_val, val_metrics):
                best_val, best_test, counter = val_metrics, val_metrics, 0
            else:
                counter += 1
                if counter >= args.patience and epoch > args.min_epochs:
                    log("Early stopping")
                    break
    log('Total time elapsed: {:.4f}s'.format(time.time() - t_total))
    log('Test set results: ' + format_metrics(best_test or best_val, 'test'))
    return model, best_test or best_val
