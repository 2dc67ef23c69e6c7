"""Pin the model oracle to REAL TensorFlow (run on any machine with tensorflow >= 2.13; none is installable in the build
container, so the committed model goldens come from oracle/tfgraph_interp.py and the oracle header says "parity unpinned").

For each released graph this runs exactly what the reference runs (src/download_and_predict_job.py:353-357,
`sess.run(predict_logits, feed_dict={predict_inp: x, predict_length: lengths})` on `conv2d/Sigmoid:0`, :1819) on the seeded
input of sentinel_tree_cover_b200.synth / oracle.preproc_ref.synth_model_input and writes

    tests/golden/model_tf_<size>.npz    y, seed, batch, length, and the graph's 60 weight tensors under "w/<name>"
    tests/golden/superresolve_tf.npz    x, y of superresolve_graph.pb (:115-117)

tests/test_oracle_model.py::test_restatement_matches_real_tensorflow_golden_when_present and
tests/test_gpu_model.py::test_gpu_matches_real_tensorflow_golden_when_present pick the files up when they exist.

Usage: python tools/make_golden_tf.py /path/to/sentinel-tree-cover [--sizes 76,124,172,220]"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    sizes = [76, 124, 172, 220]
    if "--sizes" in sys.argv:
        sizes = [int(v) for v in sys.argv[sys.argv.index("--sizes") + 1].split(",")]
    import tensorflow as tf                                    # fails loudly where TF is absent
    tf1 = tf.compat.v1
    tf1.disable_eager_execution()
    from oracle import preproc_ref as P
    from sentinel_tree_cover_b200.weights import load_predict_pb
    gold = os.path.join(ROOT, "tests", "golden")
    for size in sizes:
        pb = os.path.join(ref, "models-release", "master-ckpt-frozen", "predict_graph-%d.pb" % size)
        gd = tf1.GraphDef()
        gd.ParseFromString(open(pb, "rb").read())
        g = tf.Graph()
        with g.as_default():
            tf1.import_graph_def(gd, name="predict")              # the reference's scope name (:1789)
        sess = tf1.Session(graph=g)
        seed, batch = 100 + size, 2
        x = P.synth_model_input(batch, size, seed)
        length = np.array([4, 3], np.int64)
        y = sess.run(g.get_tensor_by_name("predict/conv2d/Sigmoid:0"),
                     feed_dict={g.get_tensor_by_name("predict/Placeholder:0"): x,
                                g.get_tensor_by_name("predict/PlaceholderWithDefault:0"): length})
        out = {"y": np.asarray(y)[..., 0].astype(np.float32), "seed": seed, "batch": batch, "length": length,
               "tf_version": np.array(tf.__version__)}
        for k, v in load_predict_pb(pb).items():
            out["w/" + k] = v
        np.savez_compressed(os.path.join(gold, "model_tf_%d.npz" % size), **out)
        print("wrote model_tf_%d.npz" % size, out["y"].shape, float(out["y"].mean()))
    pb = os.path.join(ref, "models-release", "supres-40k-swir", "superresolve_graph.pb")
    gd = tf1.GraphDef()
    gd.ParseFromString(open(pb, "rb").read())
    g = tf.Graph()
    with g.as_default():
        tf1.import_graph_def(gd, name="superresolve")
    sess = tf1.Session(graph=g)
    r = np.random.default_rng(77)
    x = r.uniform(0.0, 0.5, (3, 48, 44, 10)).astype(np.float32)
    y = sess.run(g.get_tensor_by_name("superresolve/Add_2:0"),
                 feed_dict={g.get_tensor_by_name("superresolve/Placeholder:0"): x,
                            g.get_tensor_by_name("superresolve/Placeholder_1:0"): x[..., 4:]})
    np.savez_compressed(os.path.join(gold, "superresolve_tf.npz"), x=x, y=np.asarray(y, np.float32), tf_version=np.array(tf.__version__))
    print("wrote superresolve_tf.npz")


if __name__ == "__main__":
    main()
