"""Static check of the programmatic-dependent-launch contract (csrc/kernels_fp32.cuh): every kernel that espnet_api.cu launches
through launch_k -- i.e. possibly with cudaLaunchAttributeProgrammaticStreamSerialization -- must call pdl_trigger() and
pdl_wait(), and must not touch global activations before the wait.  A kernel launched that way WITHOUT the wait would start
reading its input while the kernel before it is still writing it."""
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "glomeruli_segmentation_b200", "csrc")


def _kernel_bodies():
    bodies = {}
    for f in os.listdir(CSRC):
        if not f.endswith(".cuh"):
            continue
        src = open(os.path.join(CSRC, f)).read()
        for m in re.finditer(r"__global__\s+void\s+(?:__launch_bounds__\([^)]*\)\s*)?(\w+)\s*\(", src):
            start = src.index("{", m.end())
            depth, i = 0, start
            while True:
                depth += {"{": 1, "}": -1}.get(src[i], 0)
                if depth == 0:
                    break
                i += 1
            bodies[m.group(1)] = src[start:i + 1]
    return bodies


def test_every_kernel_launched_with_launch_k_waits_for_its_predecessor():
    api = open(os.path.join(CSRC, "espnet_api.cu")).read()
    names = set(re.findall(r"launch_k\(h,\s*kPdl\w+,\s*(\w+)", api))
    assert "kern" in names                      # the branch kernel goes through a function-pointer variable
    names.discard("kern")
    names.add("esp_branch_tc_kernel")
    assert {"stem_kernel", "reduce1x1_tc_kernel", "reduce3x3s2_tma_kernel", "reduce3x3s2_tc_kernel", "head3v_kernel",
            "dec_av_kernel", "dec_b4_kernel", "dec_c4_kernel", "upsample8_argmax_kernel"} <= names
    bodies = _kernel_bodies()
    for n in sorted(names):
        body = bodies[n]
        assert "pdl_trigger();" in body and "pdl_wait();" in body, n
        before = body[:body.index("pdl_wait();")]
        assert body.index("pdl_trigger();") < body.index("pdl_wait();"), n
        # nothing written to global memory and no TMA / bulk load of activations before the wait
        assert "tma_load_4d" not in before and "bulk_g2s" not in before, n
        assert not re.search(r"\bp\.(out\w*|logits|mask|prob_acc|enc_out|up_out|tout|comb)\s*(\[|\+)", before), n


def test_kernels_launched_without_the_attribute_are_not_in_the_chain_by_accident():
    """<<< >>> launches in espnet_api.cu are plain stream order; the helper is the only place that sets the attribute."""
    api = open(os.path.join(CSRC, "espnet_api.cu")).read()
    assert api.count("cudaLaunchAttributeProgrammaticStreamSerialization") == 1
    assert "cudaLaunchKernelEx" in api[api.index("cudaError_t launch_k("):api.index("cudaError_t launch_k(") + 1200]
