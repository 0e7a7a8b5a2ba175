#!/usr/bin/env python3
"""Build the reference-derived checkers into oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (present in the build
container, absent on the GPU box -- the GPU box only uses the prebuilt files,
which travel with the snapshot).  Reference sources are compiled where they
lie; the few that need a patch are patched in a temporary directory outside the
repository and never written into the tree.  Outputs:

  oracle/_ref/libref_stream.so  the reference's own header-only scene classes
        (Vec3.h .. Camera.h) compiled for the HOST by g++ through the shims in
        oracle/shim/, plus CreateWorld (kernel.cu:157-545) as text, driven by
        oracle/ref_stream_main.cpp.  Random numbers come from the render path's
        counter-based stream.  Built -O2 -ffp-contract=off: this is what pins
        oracle/rt_oracle.cpp bit for bit.
  oracle/_ref/libref_stream_fast.so  the same sources built -O3 -march=native:
        the CPU baseline ("kind": "reference") that bench.py times.
  oracle/_ref/ref_gpu           the reference's kernel.cu for sm_100 with its
        hard-coded constants turned into argv, a ray counter and cudaEvent
        timing: the GPU baseline the >=10x target is measured against, and the
        statistical image reference (unmodified FP64 arithmetic + cuRAND XORWOW).
  oracle/_ref/earthmap.jpg      the reference's texture asset, for ref_gpu.

Patches (each must match exactly once, or the build fails):
  T1  sequencing of multi-draw expressions g++ would evaluate right-to-left but
      nvcc device code evaluates left-to-right: Material.h:19-21, Camera.h:15,
      Perlin.h:91-93, kernel.cu:216,229,237-238,502.
  T3  ConstantMedium.h:79 draws from the keyed medium stream; the medium gets an
      id in construction order.
  B1  scene id 10 = "Book 1 final": scene 0 (kernel.cu:199-258) with static
      spheres, grey ground, closed shutter; same RNG draws as scene 0.
  ref_gpu only: argv for size/scene/spp/seed, ray counter, event timing,
      optional raw framebuffer dump, optional scene-box dump.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RT_REFERENCE_DIR", "/root/reference")
RT = os.path.join(REF, "RayTracinginOneWeekend")
OUT = os.path.join(HERE, "_ref")

HEADERS = [
    "Vec3.h", "Interval.h", "Ray.h", "AABB.h", "Hittable.h", "HittableList.h", "BvhNode.h", "Sphere.h",
    "MovingSphere.h", "Quad.h", "Instance.h", "ConstantMedium.h", "Texture.h", "Perlin.h", "Material.h",
    "Metal.h", "Dielectric.h", "Camera.h",
]


def sub_once(text: str, pattern: str, repl: str, what: str, flags=re.S) -> str:
    new, n = re.subn(pattern, lambda m: repl, text, flags=flags)
    if n != 1:
        raise SystemExit(f"build_ref: patch '{what}' matched {n} times (expected 1)")
    return new


def ws(s: str) -> str:
    """Literal text -> regex tolerant of whitespace runs."""
    return r"\s*".join(re.escape(tok) for tok in s.split())


def patched_headers(tmp: str) -> None:
    for h in HEADERS:
        with open(os.path.join(RT, h), encoding="utf-8") as f:
            src = f.read()
        if h == "Material.h":
            src = sub_once(
                src,
                ws("p = 2.0 * Vector3(curand_uniform(randState), curand_uniform(randState), "
                   "curand_uniform(randState)) - Vector3(1.0, 1.0, 1.0);"),
                "{ const double rx_ = curand_uniform(randState); const double ry_ = curand_uniform(randState); "
                "const double rz_ = curand_uniform(randState); "
                "p = 2.0 * Vector3(rx_, ry_, rz_) - Vector3(1.0, 1.0, 1.0); }",
                "T1 Material.h:19-21")
        elif h == "Camera.h":
            src = sub_once(
                src,
                ws("p = 2.0 * Vector3(curand_uniform(randState), curand_uniform(randState), 0.0) "
                   "- Vector3(1.0, 1.0, 0.0);"),
                "{ const double rx_ = curand_uniform(randState); const double ry_ = curand_uniform(randState); "
                "p = 2.0 * Vector3(rx_, ry_, 0.0) - Vector3(1.0, 1.0, 0.0); }",
                "T1 Camera.h:15")
        elif h == "Perlin.h":
            src = sub_once(
                src,
                ws("return Vector3(min + range * curand_uniform(s), min + range * curand_uniform(s), "
                   "min + range * curand_uniform(s));"),
                "{ const double rx_ = min + range * curand_uniform(s); const double ry_ = min + range * "
                "curand_uniform(s); const double rz_ = min + range * curand_uniform(s); "
                "return Vector3(rx_, ry_, rz_); }",
                "T1 Perlin.h:91-93")
        elif h == "ConstantMedium.h":
            src = sub_once(src, ws("log(curand_uniform(randState))"),
                           "log(rtshim_medium_uniform(randState, mMediumId))", "T3 ConstantMedium.h:79")
            src = sub_once(src, ws("Material* mPhaseFunction;"),
                           "Material* mPhaseFunction;\n    int mMediumId = rtshim_next_medium_id();",
                           "T3 medium id member")
        with open(os.path.join(tmp, h), "w", encoding="utf-8") as f:
            f.write(src)


def book1_variant(block: str) -> str:
    """Scene 0 block -> the scene-10 (Book 1 final) block."""
    v = sub_once(block, ws("if (sceneId == 0)"), "else if (sceneId == 10)", "B1 head")
    v = sub_once(v, ws("Vector3(0.0, -1000.0, -1.0), 1000.0, new Lambertian(checker));"),
                 "Vector3(0.0, -1000.0, -1.0), 1000.0, new Lambertian(Color(0.5, 0.5, 0.5)));", "B1 ground")
    v = sub_once(v, ws("new MovingSphere( center, center2, 0.0, 1.0, 0.2,"), "new Sphere(center, 0.2,",
                 "B1 static spheres")
    v = sub_once(v, ws("shutterClose = 1.0;"), "shutterClose = 0.0;", "B1 shutter")
    return v


def add_book1(src: str) -> str:
    m = re.search(r"if \(sceneId == 0\).*?(?=else if \(sceneId == 1\))", src, flags=re.S)
    if not m:
        raise SystemExit("build_ref: scene 0 block not found")
    block = m.group(0)
    return src[:m.end()] + book1_variant(block) + "\t\t" + src[m.end():]


def create_world_text(for_host: bool) -> str:
    with open(os.path.join(RT, "kernel.cu"), encoding="utf-8") as f:
        src = f.read()
    m = re.search(r"#define RND .*?#undef RND\n", src, flags=re.S)
    if not m:
        raise SystemExit("build_ref: CreateWorld span not found")
    cw = add_book1(m.group(0))
    if for_host:
        cw = cw.replace("Vector3 center(a + 0.9 * RND, 0.2, b + 0.9 * RND);",
                        "Vector3 center = RtShimCenter(&localRandState, a, b);")
        if cw.count("RtShimCenter(") != 2:
            raise SystemExit("build_ref: T1 kernel.cu:216 patch count")
        cw = cw.replace("new Lambertian(Color(RND * RND, RND * RND, RND * RND))",
                        "new Lambertian(RtShimColorProducts(&localRandState))")
        if cw.count("RtShimColorProducts(") != 2:
            raise SystemExit("build_ref: T1 kernel.cu:229 patch count")
        cw, n = re.subn(ws("new Metal( Color(0.5 * (1.0 + RND), 0.5 * (1.0 + RND), 0.5 * (1.0 + RND)), 0.5 * RND)"),
                        lambda _m: "RtShimNewMetal(&localRandState)", cw)
        if n != 2:
            raise SystemExit("build_ref: T1 kernel.cu:237 patch count")
        cw = sub_once(cw, ws("Point3 c(165.0 * RND, 165.0 * RND, 165.0 * RND);"),
                      "Point3 c = RtShimPoint(&localRandState, 165.0);", "T1 kernel.cu:502")
    return cw


def run(cmd: list[str], cwd: str | None = None) -> None:
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=cwd)


def build_ref_stream(tmp: str) -> None:
    patched_headers(tmp)
    with open(os.path.join(tmp, "create_world.inc"), "w", encoding="utf-8") as f:
        f.write(create_world_text(for_host=True))
    # The reference's stb translation unit, compiled where it lies.
    run(["g++", "-O2", "-fPIC", "-w", "-c", os.path.join(RT, "StbImageImpl.cpp"), "-o",
         os.path.join(tmp, "stb.o")])
    run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-w",
         "-I", tmp, "-I", os.path.join(HERE, "shim"),
         os.path.join(HERE, "ref_stream_main.cpp"), os.path.join(HERE, "ref_image.cpp"),
         os.path.join(tmp, "stb.o"), "-o", os.path.join(OUT, "libref_stream.so")])
    # The same sources as an optimised build: the CPU BASELINE of bench.py (-O3 -march=native).  Not a
    # pin: contraction and vectorisation may change last bits; tests/test_oracle_pin.py checks it stays within 1e-9.
    run(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w",
         "-I", tmp, "-I", os.path.join(HERE, "shim"),
         os.path.join(HERE, "ref_stream_main.cpp"), os.path.join(HERE, "ref_image.cpp"),
         os.path.join(tmp, "stb.o"), "-o", os.path.join(OUT, "libref_stream_fast.so")])


GPU_PROLOGUE = r'''
// ---- added by oracle/build_ref.py (not reference code) ----
__global__ void RtDumpBoxes(Hittable** list, int n)
{
    for (int i = 0; i < n; i++) {
        Aabb b = list[i]->BoundingBox();
        printf("BOX %d %llx %llx %llx %llx %llx %llx\n", i,
               (unsigned long long)__double_as_longlong(b.X.Min), (unsigned long long)__double_as_longlong(b.X.Max),
               (unsigned long long)__double_as_longlong(b.Y.Min), (unsigned long long)__double_as_longlong(b.Y.Max),
               (unsigned long long)__double_as_longlong(b.Z.Min), (unsigned long long)__double_as_longlong(b.Z.Max));
    }
}
// ---- end added ----
'''


def gpu_source() -> str:
    with open(os.path.join(RT, "kernel.cu"), encoding="utf-8") as f:
        src = f.read()
    m = re.search(r"#define RND .*?#undef RND\n", src, flags=re.S)
    src = src[:m.start()] + add_book1(m.group(0)) + src[m.end():]
    # ray counter through RayColor
    src = sub_once(src, ws("Hittable** world, curandState* randState) { Ray currentRay = r;"),
                   "Hittable** world, curandState* randState, unsigned long long& nRays)\n{\n\tRay currentRay = r;",
                   "gpu RayColor signature")
    src = sub_once(src, ws("HitRecord rec; if (!(*world)->Hit(currentRay, 0.001, DBL_MAX, rec, randState))"),
                   "HitRecord rec;\n\t\tnRays++;\n\t\tif (!(*world)->Hit(currentRay, 0.001, DBL_MAX, rec, randState))",
                   "gpu ray count")
    src = sub_once(src, ws("col += RayColor(r, background, world, &localRandState);"),
                   "col += RayColor(r, background, world, &localRandState, nRays);", "gpu RayColor call")
    src = sub_once(src, ws("Color col(0.0, 0.0, 0.0); // 장면 배경색"),
                   "Color col(0.0, 0.0, 0.0);\n\tunsigned long long nRays = 0ULL;\n\t// 장면 배경색", "gpu nRays decl")
    src = sub_once(src, ws("randState[pixelIndex] = localRandState; col = col / double(numSamples);"),
                   "randState[pixelIndex] = localRandState;\n\tatomicAdd(&gRtRayCount, nRays);\n"
                   "\tcol = col / double(numSamples);", "gpu ray atomic")
    # render seed
    src = sub_once(src, ws("__global__ void RenderInit(int maxX, int maxY, curandState* randState)"),
                   "__global__ void RenderInit(int maxX, int maxY, curandState* randState, unsigned long long seed)",
                   "gpu RenderInit signature")
    src = sub_once(src, ws("curand_init(1984, pixelIndex, 0, &randState[pixelIndex]);"),
                   "curand_init(seed, pixelIndex, 0, &randState[pixelIndex]);", "gpu seed use")
    src = sub_once(src, ws("RenderInit<<<blocks, threads>>>(imageWidth, imageHeight, randState);"),
                   "cudaEvent_t ev0, ev1, ev2;\n\tcudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventCreate(&ev2);\n"
                   "\tcudaEventRecord(ev0);\n"
                   "\tRenderInit<<<blocks, threads>>>(imageWidth, imageHeight, randState, renderSeed);\n"
                   "\tcudaEventRecord(ev1);", "gpu RenderInit call")
    src = sub_once(src, ws("numSamples, camera, world, randState); checkCudaErrors(cudaGetLastError());"),
                   "numSamples, camera, world, randState);\n\tcudaEventRecord(ev2);\n"
                   "\tcheckCudaErrors(cudaGetLastError());", "gpu Render event")
    # argv
    src = sub_once(src, ws("int main() {"), GPU_PROLOGUE + "int main(int argc, char** argv)\n{", "gpu main")
    src = sub_once(src, ws("int imageWidth = 1440;"), "int imageWidth = (argc > 1) ? atoi(argv[1]) : 1440;", "gpu W")
    src = sub_once(src, ws("int imageHeight = 720;"), "int imageHeight = (argc > 2) ? atoi(argv[2]) : 720;", "gpu H")
    src = sub_once(src, ws("int sceneId = 9;"), "int sceneId = (argc > 3) ? atoi(argv[3]) : 9;", "gpu scene")
    src = sub_once(src, r"int numSamples = \(sceneId == 9\)[^;]*;",
                   "int numSamples = (argc > 4) ? atoi(argv[4]) : 10;\n"
                   "\tunsigned long long renderSeed = (argc > 5) ? strtoull(argv[5], 0, 10) : 1984ULL;\n"
                   "\tconst char* rawPath = (argc > 6 && argv[6][0] != '-') ? argv[6] : 0;\n"
                   "\tconst char* ppmPath = (argc > 7 && argv[7][0] != '-') ? argv[7] : 0;", "gpu spp")
    src = sub_once(src, ws("int numNodes = *d_numNodes;"),
                   "int numNodes = *d_numNodes;\n"
                   "\tif (getenv(\"RT_DUMP_SCENE\")) { RtDumpBoxes<<<1, 1>>>(list, numHittables); "
                   "checkCudaErrors(cudaDeviceSynchronize()); }", "gpu dump")
    src = sub_once(src, ws('std::cerr << "took " << timerSeconds << " seconds.\\n";'),
                   'std::cerr << "took " << timerSeconds << " seconds.\\n";\n'
                   "\t{\n\t\tfloat msInit = 0.f, msRender = 0.f;\n"
                   "\t\tcudaEventElapsedTime(&msInit, ev0, ev1); cudaEventElapsedTime(&msRender, ev1, ev2);\n"
                   "\t\tunsigned long long rays = 0ULL;\n"
                   "\t\tcudaMemcpyFromSymbol(&rays, gRtRayCount, sizeof rays);\n"
                   "\t\tprintf(\"{\\\"impl\\\": \\\"ref_gpu\\\", \\\"scene\\\": %d, \\\"width\\\": %d, \\\"height\\\": %d, "
                   "\\\"spp\\\": %d, \\\"seed\\\": %llu, \\\"render_init_ms\\\": %.3f, \\\"render_ms\\\": %.3f, "
                   "\\\"rays\\\": %llu, \\\"mrays_per_s\\\": %.3f, \\\"objects\\\": %d, \\\"nodes\\\": %d}\\n\",\n"
                   "\t\t\tsceneId, imageWidth, imageHeight, numSamples, renderSeed, msInit, msRender, rays,\n"
                   "\t\t\t(double)rays / (msRender * 1e3), numHittables, numNodes);\n"
                   "\t\tif (rawPath) { FILE* fp = fopen(rawPath, \"wb\"); if (fp) { fwrite(frameBuffer, "
                   "sizeof(Vector3), (size_t)numPixels, fp); fclose(fp); } }\n\t}", "gpu json")
    src = sub_once(src, ws('std::ofstream outFile("output.ppm");'),
                   'std::ofstream outFile(ppmPath ? ppmPath : "/dev/null");', "gpu ppm path")
    src = sub_once(src, ws("for (int j = imageHeight - 1; j >= 0; j--) { std::cerr"),
                   "for (int j = ppmPath ? imageHeight - 1 : -1; j >= 0; j--)\n\t{\n\t\tif (0) std::cerr",
                   "gpu ppm loop")
    return ("#include <cstdio>\n#include <cstdlib>\n"
            "__device__ unsigned long long gRtRayCount = 0ULL; // added by oracle/build_ref.py\n" + src)


def build_ref_gpu(tmp: str) -> None:
    # its own directory: kernel.cu's quote-includes must find the UNPATCHED headers via -I
    tmp = os.path.join(tmp, "gpu")
    os.makedirs(tmp, exist_ok=True)
    with open(os.path.join(tmp, "ref_gpu.cu"), "w", encoding="utf-8") as f:
        f.write(gpu_source())
    run(["g++", "-O2", "-w", "-c", os.path.join(RT, "StbImageImpl.cpp"), "-o", os.path.join(tmp, "stb_exe.o")])
    run(["nvcc", "-std=c++17", "-O3", "-arch=sm_100", "-w", "-I", RT, os.path.join(tmp, "ref_gpu.cu"),
         os.path.join(tmp, "stb_exe.o"), "-o", os.path.join(OUT, "ref_gpu")])
    shutil.copyfile(os.path.join(RT, "earthmap.jpg"), os.path.join(OUT, "earthmap.jpg"))


def main() -> int:
    if not os.path.isdir(RT):
        print(f"build_ref: {RT} not found -- keeping whatever is prebuilt in {OUT}")
        return 0
    os.makedirs(OUT, exist_ok=True)
    what = set(sys.argv[1:]) or {"stream", "gpu"}
    tmp = tempfile.mkdtemp(prefix="rt_ref_build_")
    try:
        if "stream" in what:
            build_ref_stream(tmp)
        if "gpu" in what:
            build_ref_gpu(tmp)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
