"""Generate tests/golden/superresolve_cv.npz: outputs of the RELEASED frozen graph
/root/reference/models-release/supres-40k-swir/superresolve_graph.pb (fetched as `Add_2:0` from `Placeholder`,
`Placeholder_1`, src/download_and_predict_job.py:115-117) executed by a THIRD-PARTY runtime: OpenCV's DNN module
(`cv2.dnn.readNetFromTensorflow`, which parses the GraphDef and runs its own convolution / MirrorPad / Tanh layers).
TensorFlow is not installed here; this is the one independent executor of a released graph this image has.  It pins the
oracle's GraphDef interpreter (oracle/tfgraph_interp.py: Conv2D weight layout, SAME / MirrorPad REFLECT borders, the 0.05
residual scaling, the bilinear skip) and the restatement (oracle/model_ref.py: SuperresolveRef) to an implementation that
shares no code with either.  The ConvGRU / U-Net graphs (predict_graph-*.pb) do not load in OpenCV (5-D reshapes of the
group normalisation, TensorArray loops): they stay pinned to the interpreter only.
Usage: python tools/make_golden_cv.py   (this container only: needs /root/reference and cv2)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PBSR = "/root/reference/models-release/supres-40k-swir/superresolve_graph.pb"


def run_opencv(pb, x):
    """x [B, H, W, 10] float32 (bands 4: bilinearly upsampled) -> Add_2 [B, H, W, 6], one image per forward (NCHW blobs)."""
    import cv2
    net = cv2.dnn.readNetFromTensorflow(pb)
    out = []
    for i in range(x.shape[0]):
        net.setInput(np.ascontiguousarray(x[i:i + 1].transpose(0, 3, 1, 2)), "Placeholder")
        net.setInput(np.ascontiguousarray(x[i:i + 1, ..., 4:].transpose(0, 3, 1, 2)), "Placeholder_1")
        out.append(net.forward("Add_2").transpose(0, 2, 3, 1).copy())
    return np.concatenate(out).astype(np.float32)


def window_input(seed, B=1, S=118):
    """A reflectance-like padded window of superresolve_large_tile (110 + 2 x 4 px, :105-112), smooth fields + texture."""
    r = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:S, 0:S] / float(S)
    base = np.stack([0.08 + 0.05 * np.sin(6.0 * xx + c) * np.cos(5.0 * yy - 0.7 * c) + 0.02 * c / 10 for c in range(10)], -1)
    x = base[None] + 0.01 * r.standard_normal((B, S, S, 10))
    return np.clip(x, 0, 1).astype(np.float32)


def main():
    import cv2
    g = np.load(os.path.join(ROOT, "tests", "golden", "superresolve.npz"))
    out = {"y_small": run_opencv(PBSR, g["x"]), "seed_window": np.array(31), "cv_version": np.array(cv2.__version__)}
    out["y_window"] = run_opencv(PBSR, window_input(31))
    print("OpenCV", cv2.__version__, "vs the interpreter-made golden:", float(np.abs(out["y_small"] - g["y"]).max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "superresolve_cv.npz"), **out)


if __name__ == "__main__":
    main()
