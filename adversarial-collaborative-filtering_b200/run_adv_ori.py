"""``run_adv_ori.py`` surface (run_adv_ori.py:17-221): the driver that produced the reference's published logs.

    python -m apr_b200.run_adv_ori --model apr --dataset Video --eval_mode all --epochs 40 --adv_epoch 20 ...

Same flags and defaults; ``--model bpr`` and ``--model apr`` are the path built here (other model names belong to
subsystems that are out of scope and are rejected).  Extra flags: --seed, --quirk/--no-quirk (Dataset.py trainList
cursor quirk).
"""
import argparse
from time import localtime, strftime

from .APR import MF, training
from .Dataset import OriginalDataset
from .utils import init_logging, write2file


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Run AMF.")
    parser.add_argument('--path', nargs='?', default='', help='Input data path.')
    parser.add_argument('--opath', nargs='?', default='aaa/', help='Output path.')
    parser.add_argument('--dataset', nargs='?', default='fsq11-sort', help='Choose a dataset.')
    parser.add_argument('--model', type=str, help='Model Name', default="pop")
    parser.add_argument('--verbose', type=int, default=1, help='Evaluate per X epochs.')
    parser.add_argument('--batch_size', type=int, default=512, help='batch_size')
    parser.add_argument('--epochs', type=int, default=10, help='Number of epochs.')
    parser.add_argument('--adv_epoch', type=int, default=0,
                        help='Add APR in epoch X, when adv_epoch is 0, it\'s equivalent to pure AMF.\n '
                             'And when adv_epoch is larger than epochs, it\'s equivalent to pure MF model. ')
    parser.add_argument('--embed_size', type=int, default=64, help='Embedding size.')
    parser.add_argument('--dns', type=int, default=1, help='number of negative sample for each positive in dns.')
    parser.add_argument('--reg', type=float, default=0, help='Regularization for user and item embeddings.')
    parser.add_argument('--lr', type=float, default=0.05, help='Learning rate.')
    parser.add_argument('--reg_adv', type=float, default=1, help='Regularization for adversarial loss')
    parser.add_argument('--restore', type=str, default=None, help='The restore time_stamp for weights in \\Pretrain')
    parser.add_argument('--ckpt', type=int, default=10, help='Save the model per X epochs.')
    parser.add_argument('--task', nargs='?', default='', help='Add the task name for launching experiments')
    parser.add_argument('--adv', nargs='?', default='grad',
                        help='Generate the adversarial sample by gradient method or random method')
    parser.add_argument('--eps', type=float, default=0.5, help='Epsilon for adversarial weights.')
    parser.add_argument('--eps_dense', type=float, default=0.5, help='Epsilon for adversarial weights.')
    parser.add_argument('--eps_conv', type=float, default=0.5, help='Epsilon for adversarial weights.')
    parser.add_argument('--eps_pos', type=float, default=0.5, help='Epsilon for adversarial weights.')
    parser.add_argument('--eval_mode', type=str, default="sample", help='Eval mode: sample or all')
    parser.add_argument('--seed', type=int, default=2019, help='Philox seed of init and sampler (extra flag).')
    parser.add_argument('--no-quirk', dest='quirk', action='store_false',
                        help='Group trainList by true uid instead of reproducing Dataset.py:316-320 (extra flag).')
    return parser.parse_args(argv)


def main(argv=None):
    time_stamp = strftime('%Y_%m_%d_%H_%M_%S', localtime())
    args = parse_args(argv)
    init_logging(args, time_stamp)
    dataset = OriginalDataset(args.path + "data/" + args.dataset, reproduce_quirk=args.quirk)
    out = args.path + "out/" + args.opath

    if args.model == "bpr":
        runName = "%s_%s_d%d_%s" % (args.dataset, args.model, args.embed_size, time_stamp)
        write2file(out, runName + ".out", runName)
        args.adver = 0
        MF_BPR = MF(dataset.num_users, dataset.num_items, args)
        MF_BPR.build_graph()
        write2file(out, runName + ".out", "Initialize MF_BPR")
        return training(MF_BPR, dataset, args, runName, epoch_start=0, epoch_end=args.epochs, time_stamp=time_stamp)

    if args.model == "apr":
        runName = "%s_%s_d%d_e%f_l%f_%s" % (args.dataset, args.model, args.embed_size, args.eps, args.reg_adv, time_stamp)
        write2file(out, runName + ".out", runName)
        args.adver = 0
        MF_BPR = MF(dataset.num_users, dataset.num_items, args)
        MF_BPR.build_graph()
        write2file(out, runName + ".out", "Initialize BPR")
        training(MF_BPR, dataset, args, runName, epoch_start=0, epoch_end=args.adv_epoch - 1, time_stamp=time_stamp)
        args.adver = 1
        AMF = MF(dataset.num_users, dataset.num_items, args)
        AMF.build_graph()
        write2file(out, runName + ".out", "Initialize APR")
        return training(AMF, dataset, args, runName, epoch_start=args.adv_epoch, epoch_end=args.epochs,
                        time_stamp=time_stamp)

    raise SystemExit("--model %s is outside the APR/BPR-MF hot path built here (use bpr or apr)" % args.model)


if __name__ == '__main__':
    main()
