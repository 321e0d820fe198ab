"""Runs the REFERENCE's own, unmodified training flow -- main.py's sequence BaseOptions().get_arguments() ->
set_seed -> trainer_node_classification.trainer(args, seed) -> trainer.main() (trainer_node_classification.py:8-29,
252-369) -- with either

  --arm ours       this repo's GNN_model ahead of the reference's on sys.path (the drop-in, CUDA kernels), or
  --arm reference  the reference's own GNN_model on the CPU, its DGL calls served by shims/dgl (index_add_),

and writes the per-epoch training loss / accuracies the trainer produced plus the class that was instantiated.
TEST INFRASTRUCTURE (tests/test_dropin_trainer.py).  Nothing of the reference is edited or copied: its directory is
put on sys.path, the wheels it imports that are absent from this image come from shims/ (import surface only), and
the synthetic "Cora" comes from shims/torch_geometric/datasets.py because the real dataset is not on disk.

The one knob the harness turns after the reference parsed its options is ``--dropout-off``: base_options.py:187-220
hard-sets dropout per dataset (0.6 for Cora) whatever the command line says, and two dropout streams (CPU Philox vs
CUDA Philox) cannot be compared; with it the harness sets args.dropout = 0 before the trainer is built.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--ref', required=True, help='reference checkout')
    p.add_argument('--arm', required=True, choices=['ours', 'reference'])
    p.add_argument('--out', required=True)
    p.add_argument('--workdir', required=True, help='the trainer writes saved_models/, figures, .npy here')
    p.add_argument('--dropout-off', action='store_true')
    p.add_argument('--cpu', action='store_true')
    p.add_argument('ref_argv', nargs='*', help='arguments for the reference option parser (after --)')
    a = p.parse_args()

    paths = [os.path.join(ROOT, 'shims'), a.ref]
    if a.arm == 'ours':
        paths.insert(0, os.path.join(ROOT, 'gnn_tail_generalization_b200'))    # shadows the reference's GNN_model/
    sys.path[:0] = paths
    os.makedirs(a.workdir, exist_ok=True)
    os.chdir(a.workdir)
    sys.argv = ['main.py'] + a.ref_argv

    import torch
    import main as ref_main                     # the reference's main.py (only its set_seed is used below)
    from base_options import BaseOptions
    args = BaseOptions().get_arguments()
    if a.cpu:
        args.cuda = False
    if a.dropout_off:
        args.dropout = 0.0
    from trainer_node_classification import trainer
    import GNN_model.GNN_normalizations as gm
    args.random_seed = 0
    ref_main.set_seed(args)
    trnr = trainer(args, 0)

    losses = []
    inner = trnr.run_trainSet

    def recording_run_trainSet():                # observes the value the trainer computed; changes nothing
        out = inner()
        losses.append(float(out[0]))
        return out
    trnr.run_trainSet = recording_run_trainSet

    launches0 = 0
    if a.arm == 'ours':
        from gnn_tail_generalization_b200 import _cabi
        launches0 = _cabi.launch_count()
    res = trnr.main()
    rec = {'arm': a.arm, 'teacher_class_file': os.path.abspath(gm.__file__),
           'teacher_class': f'{type(trnr.teacherGNN).__module__}.{type(trnr.teacherGNN).__name__}',
           'device': str(args.device), 'type_trick': args.type_trick, 'dropout': args.dropout,
           'whetherHasSE': args.whetherHasSE, 'epochs': args.epochs, 'train_loss_per_epoch': losses,
           'results_last_epoch': [float(v) for v in res[:, -1]],
           'state_dict_keys': sorted(trnr.teacherGNN.state_dict().keys()),
           'checkpoint_written': os.path.exists(os.path.join(trnr.modeldir, 'teacherGNN'))}
    if a.arm == 'ours':
        rec['kernel_launches'] = _cabi.launch_count() - launches0
        rec['cuda'] = torch.cuda.is_available()
    with open(a.out, 'w') as f:
        json.dump(rec, f)
    print(json.dumps({k: v for k, v in rec.items() if k != 'state_dict_keys'}))


if __name__ == '__main__':
    main()
