"""PyTorch-idiom port of the reference's 2-D synthetic flow trainer (run_grevnet.py) on top of the
B200 hot path.  Same flag names and defaults for everything that reaches the hot path or the
optimiser (run_grevnet.py:39-132); TF-1 session / summary / checkpoint plumbing is replaced by a
plain loop, python logging and torch.save.

    python -m graph_normalizing_flows_b200.run_grevnet --dataset mog_4 --num_train_iters 200
    python -m graph_normalizing_flows_b200.run_grevnet --make_gnn_fn sum_concat_then_mlp --use_batch_norm=False

Differences, stated: `--make_gnn_fn` accepts the reference default dm_self_attn (SURVEY §8 row f1, fp32 kernels)
and the four message-passing factories on the fused tcgen05 path (avg_then_mlp, sum_then_mlp,
avg_concat_then_mlp, sum_concat_then_mlp); the dense-attention / GRU factories of run_grevnet.py:254-266 are
out of scope.  `--use_gnf` / GNFBlock does not exist in the reference tree (grevnet.py is missing) -- GRevNet is
always used and its backward is already the reversible one.  With --use_batch_norm (reference default True)
the batch-norm bijector's gamma/beta are trained too; the reference's gamma constraint relu(gamma)+1e-6
(gnn.py:261-262, a Keras constraint) is applied as a projection after every optimiser step.
"""
from __future__ import annotations

import logging
import os
import time
from functools import partial

import numpy as np
import torch
from absl import app, flags

from . import gnn, loss as loss_lib
from .grevnet_synthetic_data import DATASETS_MAP

FLAGS = flags.FLAGS
# Graph params (run_grevnet.py:39-43)
flags.DEFINE_integer("node_embedding_dim", 2, "Number of dimensions in node embeddings.")
# GRevNet params (:46-52)
flags.DEFINE_integer("num_coupling_layers", 12, "Number of coupling layers in GRevNet.")
flags.DEFINE_bool("weight_sharing", False, "")
# GNN params (:55-72)
flags.DEFINE_string("make_gnn_fn", "dm_self_attn", "dm_self_attn | avg_then_mlp | sum_then_mlp | avg_concat_then_mlp | "
                    "sum_concat_then_mlp")
flags.DEFINE_integer("gnn_num_layers", 5, "Number of layers to use in MLP in GRevNet.")
flags.DEFINE_integer("gnn_latent_dim", 256, "Latent dim for GNN used in GRevNet.")
flags.DEFINE_float("gnn_bias_init_stddev", 0.1, "Used to initialize biases in GRevNet MLPs.")
flags.DEFINE_float("gnn_l2_regularizer_weight", 0.1, "Unused (as in the reference).")
flags.DEFINE_float("gnn_avg_then_mlp_epsilon", 1.0, "Weight of the node's own embedding vs. its neighbours'.")
# Attention params (:74-80)
flags.DEFINE_integer("attn_kq_dim", 10, "")
flags.DEFINE_integer("attn_v_dim", 10, "")
flags.DEFINE_integer("attn_num_heads", 8, "")
flags.DEFINE_integer("attn_concat_heads_output_dim", 80, "")
flags.DEFINE_bool("attn_concat", True, "")
flags.DEFINE_bool("attn_residual", False, "")
flags.DEFINE_bool("attn_layer_norm", False, "snt.LayerNorm on the GNN output.")
# Training params (:82-109)
flags.DEFINE_bool("use_batch_norm", True, "TFP batch-norm bijector between half steps (reference default).")
flags.DEFINE_string("dataset", "mog_4", "Which dataset to use.")
flags.DEFINE_string("logdir", "test_runs/test_grevnet", "Where to write training files.")
flags.DEFINE_integer("train_batch_size", 32, "Batch size used at training.")
flags.DEFINE_integer("num_train_iters", 15000, "Number of steps to run training.")
flags.DEFINE_integer("save_every_n_steps", 10000, "How often to save model.")
flags.DEFINE_integer("log_every_n_steps", 50, "How often to log model stats.")
flags.DEFINE_integer("random_seed", 12345, "")
flags.DEFINE_float("last_layer_init_scale", 1.0, "Extra damping of the last MLP layer at init (not in the reference).")
# Optimizer params (:115-131)
flags.DEFINE_float("lr", 1e-04, "Learning rate.")
flags.DEFINE_bool("use_lr_decay", True, "Whether to decay learning rate.")
flags.DEFINE_integer("lr_decay_steps", 1000, "How often to decay learning rate.")
flags.DEFINE_float("lr_decay_rate", 0.96, "How much to decay learning rate.")
flags.DEFINE_float("adam_beta1", 0.9, "Adam optimizer beta1.")
flags.DEFINE_float("adam_beta2", 0.9, "Adam optimizer beta2.")
flags.DEFINE_float("adam_epsilon", 1e-08, "Adam optimizer epsilon.")
flags.DEFINE_bool("clip_gradient_by_value", False, "")
flags.DEFINE_float("clip_gradient_value_lower", -1.0, "")
flags.DEFINE_float("clip_gradient_value_upper", 5.0, "")
flags.DEFINE_bool("clip_gradient_by_norm", False, "")
flags.DEFINE_float("clip_gradient_norm", 10.0, "")
flags.DEFINE_string("math", "", "tc3x | tc2x | bf16 | fp32 (default: tc3x when the shape allows)")


def make_gnn_fn_map():
    half = FLAGS.node_embedding_dim / 2          # run_grevnet.py:157: a float under true division
    mlp = partial(gnn.make_mlp_model, FLAGS.gnn_latent_dim, half, FLAGS.gnn_num_layers, gnn.leaky_relu,
                  FLAGS.gnn_l2_regularizer_weight, FLAGS.gnn_bias_init_stddev)
    relu_mlp = partial(gnn.make_mlp_model, FLAGS.gnn_latent_dim, half, FLAGS.gnn_num_layers, gnn.relu,
                       FLAGS.gnn_l2_regularizer_weight, FLAGS.gnn_bias_init_stddev)      # run_grevnet.py:203-206
    return {                                     # run_grevnet.py:254-266 (in-scope entries)
        "dm_self_attn": lambda: gnn.dm_self_attn_gnn(                                    # run_grevnet.py:199-211
            kq_dim=FLAGS.attn_kq_dim, v_dim=FLAGS.attn_v_dim, make_mlp_fn=relu_mlp, num_heads=FLAGS.attn_num_heads,
            concat_heads_output_dim=FLAGS.attn_concat_heads_output_dim, concat=FLAGS.attn_concat,
            residual=FLAGS.attn_residual, layer_norm=FLAGS.attn_layer_norm),
        "avg_then_mlp": lambda: gnn.avg_then_mlp_gnn(mlp, FLAGS.gnn_avg_then_mlp_epsilon),
        "sum_then_mlp": lambda: gnn.sum_then_mlp_gnn(mlp, FLAGS.gnn_avg_then_mlp_epsilon),
        "avg_concat_then_mlp": lambda: gnn.avg_concat_then_mlp_gnn(mlp),
        "sum_concat_then_mlp": lambda: gnn.sum_concat_then_mlp_gnn(mlp),
    }


def main(argv):
    del argv
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    torch.manual_seed(FLAGS.random_seed)
    np.random.seed(FLAGS.random_seed)
    import random
    random.seed(FLAGS.random_seed)
    logdir = os.path.join(os.environ.get("MLPATH", "."), FLAGS.logdir)        # run_grevnet.py:142-146
    os.makedirs(logdir, exist_ok=True)
    dataset = DATASETS_MAP[FLAGS.dataset]
    grevnet = gnn.GRevNet(make_gnn_fn_map()[FLAGS.make_gnn_fn], FLAGS.num_coupling_layers, FLAGS.node_embedding_dim,
                          use_batch_norm=FLAGS.use_batch_norm, weight_sharing=FLAGS.weight_sharing,
                          seed=FLAGS.random_seed, math=FLAGS.math or None)          # run_grevnet.py:277-281
    if FLAGS.last_layer_init_scale != 1.0:
        grevnet.scale_last_layers_(FLAGS.last_layer_init_scale)
    train_vars = [grevnet.params] + ([grevnet.bn_gamma, grevnet.bn_beta] if FLAGS.use_batch_norm else [])
    opt = torch.optim.Adam(train_vars, lr=FLAGS.lr, betas=(FLAGS.adam_beta1, FLAGS.adam_beta2),
                           eps=FLAGS.adam_epsilon)                                   # run_grevnet.py:357-361
    t0 = time.time()
    for step in range(1, FLAGS.num_train_iters + 1):
        if FLAGS.use_lr_decay:                                                       # tf.train.exponential_decay(staircase=False) :346-352
            for grp in opt.param_groups:
                grp["lr"] = FLAGS.lr * FLAGS.lr_decay_rate ** ((step - 1) / FLAGS.lr_decay_steps)   # global_step = step - 1
        graph = dataset.get_next_batch(FLAGS.train_batch_size).to("cuda")
        scalars, grads = grevnet.loss_and_grad(graph, per_node=False)               # total_loss, :296,362-364
        if FLAGS.clip_gradient_by_value:                                             # :365-370
            for v in train_vars:
                v.grad.clamp_(FLAGS.clip_gradient_value_lower, FLAGS.clip_gradient_value_upper)
        if FLAGS.clip_gradient_by_norm:                                              # :372-375 (per-variable in TF)
            for v in train_vars:
                torch.nn.utils.clip_grad_norm_([v], FLAGS.clip_gradient_norm)
        opt.step()
        if FLAGS.use_batch_norm:                                                     # gamma_constraint, gnn.py:261-262
            with torch.no_grad():
                grevnet.bn_gamma.copy_(torch.relu(grevnet.bn_gamma) + 1e-6)
        if step % FLAGS.log_every_n_steps == 0 or step == 1:                         # scalars of :385-392
            z = scalars["z"].nodes
            logging.info("step %d loss_per_node %.5f log_prob_zs_per_node %.5f log_det_jacobian_per_node %.5f "
                         "z mean %.3f std %.3f  (%.1f steps/s)", step, float(scalars["loss_per_node"]),
                         float(scalars["log_prob_zs_per_node"]), float(scalars["log_det_jacobian_per_node"]),
                         float(z.mean()), float(z.std()), step / (time.time() - t0))
        if step % FLAGS.save_every_n_steps == 0 or step == FLAGS.num_train_iters:
            torch.save(grevnet.state_dict(), os.path.join(logdir, f"grevnet_{step}.pt"))
    # sampling pass (run_grevnet.py:304-311): z ~ N(0, I) -> x = g(z)
    graph = dataset.get_next_batch(FLAGS.train_batch_size).to("cuda")
    sample = graph.replace(nodes=torch.randn_like(graph.nodes))
    top = grevnet(sample, inverse=False).nodes
    logging.info("samples: mean %s std %s", top.mean(0).tolist(), top.std(0).tolist())
    return 0


if __name__ == "__main__":
    app.run(main)
