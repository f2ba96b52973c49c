// Command-line front end with the reference's usage (src/main/pbrt.cpp:76-173):
//   pb2_pbrt [--nthreads n] [--outfile f] [--cropwindow x0 x1 y0 y1] [--quiet] scene.pbrt ...
// Parsing, BVH build and film output run on the host; Integrator::Render runs on the GPU.
#include <cstdlib>

#include "api.h"

using namespace pbrt;

static void usage(const char *msg = nullptr) {
    if (msg) std::fprintf(stderr, "pb2_pbrt: %s\n\n", msg);
    std::fprintf(stderr,
                 "usage: pb2_pbrt [<options>] <filename.pbrt...>\n"
                 "  --cropwindow <x0,x1,y0,y1>  Specify an image crop window.\n"
                 "  --help               Print this help text.\n"
                 "  --nthreads <num>     Accepted for compatibility (rendering runs on the GPU).\n"
                 "  --outfile <filename> Write the final image to the given filename (.exr, .pfm, .png, .tga).\n"
                 "  --quick              Automatically reduce a number of quality settings to render more quickly.\n"
                 "  --quiet              Suppress all text output other than error messages.\n");
    std::exit(msg ? 1 : 0);
}

int main(int argc, char *argv[]) {
    Options options;
    std::vector<std::string> filenames;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--nthreads" || a == "-nthreads") {
            if (i + 1 == argc) usage("missing value after --nthreads argument");
            options.nThreads = std::atoi(argv[++i]);
        } else if (a.compare(0, 11, "--nthreads=") == 0) {
            options.nThreads = std::atoi(a.c_str() + 11);
        } else if (a == "--outfile" || a == "-outfile") {
            if (i + 1 == argc) usage("missing value after --outfile argument");
            options.imageFile = argv[++i];
        } else if (a.compare(0, 10, "--outfile=") == 0) {
            options.imageFile = a.substr(10);
        } else if (a == "--cropwindow" || a == "-cropwindow") {
            if (i + 4 >= argc) usage("missing value after --cropwindow argument");
            options.cropWindow[0][0] = (Float)std::atof(argv[++i]);
            options.cropWindow[0][1] = (Float)std::atof(argv[++i]);
            options.cropWindow[1][0] = (Float)std::atof(argv[++i]);
            options.cropWindow[1][1] = (Float)std::atof(argv[++i]);
        } else if (a == "--quick" || a == "-quick") {
            options.quickRender = true;
        } else if (a == "--quiet" || a == "-quiet") {
            options.quiet = true;
        } else if (a == "--logtostderr" || a == "-logtostderr" || a == "--verbose" || a == "-verbose" || a == "--v" || a == "-v") {
            // logging switches of the reference's glog front end (pbrt.cpp:118-139): nothing to configure here
        } else if (a == "--minloglevel" || a == "-minloglevel" || a == "--logdir" || a == "-logdir") {
            if (i + 1 == argc) usage(("missing value after " + a + " argument").c_str());
            ++i;
        } else if (a.compare(0, 14, "--minloglevel=") == 0 || a.compare(0, 9, "--logdir=") == 0 || a.compare(0, 4, "--v=") == 0) {
        } else if (a == "--cat" || a == "-cat" || a == "--toply" || a == "-toply") {
            usage((a + ": scene re-printing is not part of this build").c_str());
        } else if (a == "--help" || a == "-help" || a == "-h") {
            usage();
        } else if (a.size() > 1 && a[0] == '-') {
            usage(("unknown option " + a).c_str());
        } else
            filenames.push_back(a);
    }
    if (filenames.empty()) usage("no scene file given");
    pbrtInit(options);
    for (const std::string &f : filenames) {
        pbrtParseFile(f);
        if (!options.quiet && pbrtLastSetup()) {
            PathIntegrator *pi = dynamic_cast<PathIntegrator *>(pbrtLastSetup()->integrator.get());
            if (pi) {
                const pb2_stats &st = pi->lastStats;
                std::printf("Rendered %s: %.1f ms on the device, %llu camera rays, %llu regular + %llu shadow ray tests\n",
                            f.c_str(), st.render_ms, (unsigned long long)st.camera_rays,
                            (unsigned long long)st.regular_rays, (unsigned long long)st.shadow_rays);
            }
        }
    }
    pbrtCleanup();
    return g_errorCount ? 1 : 0;
}
