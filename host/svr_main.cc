// svr_main.cc -- SVRreconstructionGPU: the reference's command line (source/reconstructionGPU2/reconstruction.cc:54-1309)
// over the B200 library.  Same option names, defaults, call order, log files and output file names as the reference's
// GPU path.  What differs (printed at start-up):
//   * there is no CPU path: --useCPU is refused (no CPU fallback by design);
//   * every registration runs on the GPU: the reference's default IRTK registrations (--useCPUReg: StackRegistrations and the
//     per-slice irtkImageRigidRegistrationWithPadding, on the CPU through TBB there) on the device engine svr_rreg_register
//     (same optimiser path, bit for bit), --useGPUReg the NCC registration of Reconstruction::registerSlicesToVolume;
//     --noStackRegistration (not a reference option) skips the volumetric stack registration for pre-aligned input;
//   * --patchBased / --superpixelBased belong to PVRreconstructionGPU and are refused here.
// The option parser restates the boost::program_options behaviour the reference relies on: multitoken options take
// every following token up to the next option; po::value<bool> options take a value; bool_switch options take none.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include "svr_reconstruction.h"

namespace {

// ---- PerfStats, include/perfstats.h:44-106 (same keys and print format) --------------------------------------------
struct PerfStats {
    std::map<std::string, std::vector<double>> stats;
    std::vector<std::string> order;
    std::chrono::steady_clock::time_point last;
    void start() { last = std::chrono::steady_clock::now(); }
    void sample(const std::string& key)
    {
        const auto now = std::chrono::steady_clock::now();
        const double t = std::chrono::duration<double>(now - last).count();
        if (!stats.count(key)) order.push_back(key);
        stats[key].push_back(t);
        last = std::chrono::steady_clock::now();
    }
    void print(std::ostream& out) const
    {
        for (const auto& it : stats) {               // std::map order, as in the reference
            out << it.first << ":";
            out << std::string("\t\t\t").substr(0, 3 - std::min<size_t>(3, (it.first.size() + 1) >> 3));
            const double sum = std::accumulate(it.second.begin(), it.second.end(), 0.0);
            const double avg = sum / std::max<size_t>(it.second.size(), 1);
            const double mx = *std::max_element(it.second.begin(), it.second.end());
            out << avg * 1000.0 << " ms" << "\t(max = " << mx * 1000 << " ms" << ")\n";
        }
    }
};

std::string currentDateTime()
{
    time_t now = time(nullptr);
    struct tm tstruct;
    char buf[80];
    localtime_r(&now, &tstruct);
    strftime(buf, sizeof(buf), "%Y-%m-%d.%H-%M-%S", &tstruct);
    return buf;
}

struct Options {
    std::string output, mask, log_prefix, tfolder, sfolder, referenceVolume, manualMask, dump_setup;
    std::vector<std::string> input, transformation;
    std::vector<double> thickness;
    std::vector<int> packages, force_exclude, devices;
    int iterations = 4, levels = 3;
    double sigma = 12.0, resolution = 0.75, average = 700, delta = 150, lambda = 0.02, lastIterLambda = 0.01, smooth_mask = 4,
           low_intensity_cutoff = 0.01;
    bool global_bias_correction = false, intensity_matching = true, debug = false, debug_gpu = false, no_log = false;
    unsigned rec_iterations_first = 4, rec_iterations_last = 13, num_stacks_tuner = 0, T1PackageSize = 0, patchSize = 64, patchStride = 32;
    bool useCPU = false, useCPUReg = true, useGPUReg = false, useAutoTemplate = false, disableBiasCorr = true, patchBased = false, noStackRegistration = false,
         superpixelBased = false, useNMI = false, saveSliceTransformations = false;
    float superpixel = 0;
};

void usage()
{
    std::cout << "Application to perform reconstruction of volumetric MRI from thick slices.\nOptions:\n"
        "  -h [ --help ]                      Print usage messages\n"
        "  -o [ --output ] arg                Name for the reconstructed volume. Nifti format.\n"
        "  -m [ --mask ] arg                  Binary mask to define the region od interest.\n"
        "  -i [ --input ] arg                 [stack_1] .. [stack_N]  The input stacks.\n"
        "  -t [ --transformation ] arg        The transformations of the input stack to template in 'dof' format. Use 'id' for identity.\n"
        "  --thickness arg                    [th_1] .. [th_N] slice thickness. [Default: twice voxel size in z direction]\n"
        "  -p [ --packages ] arg              Number of packages per stack (not supported by this build)\n"
        "  --iterations arg (=4)              Number of registration-reconstruction iterations.\n"
        "  --sigma arg (=12)                  Stdev for bias field.\n"
        "  --resolution arg (=0.75)           Isotropic resolution of the volume.\n"
        "  --multires arg (=3)                Multiresolution smooting with given number of levels.\n"
        "  --average arg (=700)               Average intensity value for stacks\n"
        "  --delta arg (=150)                 Parameter to define what is an edge.\n"
        "  --lambda arg (=0.02)               Smoothing parameter.\n"
        "  --lastIterLambda arg (=0.01)       Smoothing parameter for last iteration.\n"
        "  --smooth_mask arg (=4)             Smooth the mask to reduce artefacts of manual segmentation.\n"
        "  --global_bias_correction arg (=0)  (accepted; the reference's GPU path prints 'not implemented')\n"
        "  --low_intensity_cutoff arg (=0.01)\n"
        "  --force_exclude arg                Force exclusion of slices with these indices.\n"
        "  --no_intensity_matching arg        Switch off intensity matching.\n"
        "  --log_prefix arg                   Prefix for the log file.\n"
        "  --debug arg (=0)                   Debug mode - save intermediate results.\n"
        "  --debug_gpu                        Debug only GPU results.\n"
        "  --rec_iterations_first arg (=4)    Number of superresolution iterations\n"
        "  --rec_iterations_last arg (=13)    Number of superresolution iterations for the last iteration\n"
        "  --num_stacks_tuner arg (=0)        Use only the first x input stacks\n"
        "  --no_log arg (=0)                  Do not redirect cout and cerr to log files.\n"
        "  -d [ --devices ] arg               GPUs to use, e.g. -d 0 1 2 3: one rank per device, slices sharded, NCCL all-reduce\n"
        "  --tfolder arg                      Use existing slice-to-volume transformations to initialize the reconstruction.\n"
        "  --referenceVolume arg              Optional reference volume used as inital reconstruction.\n"
        "  --useCPU / --useCPUReg / --useGPUReg / --useAutoTemplate / --disableBiasCorrection / --useNMI\n"
        "  --saveSliceTransformations         Save slice transformations.\n"
        "  --dump_setup arg                   (this build only) write the packed slices, mask and matrices that would be uploaded\n"
        "                                     to the GPU into directory arg and exit; needs no GPU.\n";
}

// returns false on a parse error (message printed)
bool parse(int argc, char** argv, Options& o, bool& help)
{
    std::map<std::string, std::string> alias = { { "-h", "--help" }, { "-o", "--output" }, { "-m", "--mask" }, { "-i", "--input" },
        { "-t", "--transformation" }, { "-p", "--packages" }, { "-d", "--devices" }, { "-s", "--superpixel" } };
    auto is_option = [&](const char* a) {
        if (a[0] != '-' || a[1] == 0) return false;
        if (a[1] == '-') return true;
        return alias.count(a) > 0;                  // a negative number is a value, not an option
    };
    int i = 1;
    auto values = [&](std::vector<std::string>& out) { while (i + 1 < argc && !is_option(argv[i + 1])) out.push_back(argv[++i]); };
    auto one = [&](const std::string& name, std::string& out) {
        if (i + 1 >= argc) { std::cerr << "ERROR: the required argument for option '" << name << "' is missing" << std::endl; return false; }
        out = argv[++i];
        return true;
    };
    auto as_bool = [](const std::string& v) { return v == "1" || v == "true" || v == "on" || v == "yes"; };
    for (; i < argc; ++i) {
        std::string a = argv[i];
        if (alias.count(a)) a = alias[a];
        std::string v;
        std::vector<std::string> vs;
        if (a == "--help") { help = true; return true; }
        else if (a == "--output") { if (!one(a, o.output)) return false; }
        else if (a == "--mask") { if (!one(a, o.mask)) return false; }
        else if (a == "--input") values(o.input);
        else if (a == "--transformation") values(o.transformation);
        else if (a == "--thickness") { values(vs); for (auto& s : vs) o.thickness.push_back(atof(s.c_str())); }
        else if (a == "--packages") { values(vs); for (auto& s : vs) o.packages.push_back(atoi(s.c_str())); }
        else if (a == "--force_exclude") { values(vs); for (auto& s : vs) o.force_exclude.push_back(atoi(s.c_str())); }
        else if (a == "--devices") { values(vs); for (auto& s : vs) o.devices.push_back(atoi(s.c_str())); }
        else if (a == "--iterations") { if (!one(a, v)) return false; o.iterations = atoi(v.c_str()); }
        else if (a == "--sigma") { if (!one(a, v)) return false; o.sigma = atof(v.c_str()); }
        else if (a == "--resolution") { if (!one(a, v)) return false; o.resolution = atof(v.c_str()); }
        else if (a == "--multires") { if (!one(a, v)) return false; o.levels = atoi(v.c_str()); }
        else if (a == "--average") { if (!one(a, v)) return false; o.average = atof(v.c_str()); }
        else if (a == "--delta") { if (!one(a, v)) return false; o.delta = atof(v.c_str()); }
        else if (a == "--lambda") { if (!one(a, v)) return false; o.lambda = atof(v.c_str()); }
        else if (a == "--lastIterLambda") { if (!one(a, v)) return false; o.lastIterLambda = atof(v.c_str()); }
        else if (a == "--smooth_mask") { if (!one(a, v)) return false; o.smooth_mask = atof(v.c_str()); }
        else if (a == "--global_bias_correction") { if (!one(a, v)) return false; o.global_bias_correction = as_bool(v); }
        else if (a == "--low_intensity_cutoff") { if (!one(a, v)) return false; o.low_intensity_cutoff = atof(v.c_str()); }
        else if (a == "--no_intensity_matching") { if (!one(a, v)) return false; o.intensity_matching = as_bool(v); }   // sic: the reference stores the value in intensity_matching
        else if (a == "--log_prefix") { if (!one(a, o.log_prefix)) return false; }
        else if (a == "--debug") { if (!one(a, v)) return false; o.debug = as_bool(v); }
        else if (a == "--debug_gpu") o.debug_gpu = true;
        else if (a == "--rec_iterations_first") { if (!one(a, v)) return false; o.rec_iterations_first = (unsigned)atoi(v.c_str()); }
        else if (a == "--rec_iterations_last") { if (!one(a, v)) return false; o.rec_iterations_last = (unsigned)atoi(v.c_str()); }
        else if (a == "--num_stacks_tuner") { if (!one(a, v)) return false; o.num_stacks_tuner = (unsigned)atoi(v.c_str()); }
        else if (a == "--no_log") { if (!one(a, v)) return false; o.no_log = as_bool(v); }
        else if (a == "--tfolder") { if (!one(a, o.tfolder)) return false; }
        else if (a == "--sfolder") { if (!one(a, o.sfolder)) return false; }
        else if (a == "--referenceVolume") { if (!one(a, o.referenceVolume)) return false; }
        else if (a == "--T1PackageSize") { if (!one(a, v)) return false; o.T1PackageSize = (unsigned)atoi(v.c_str()); }
        else if (a == "--useCPU") o.useCPU = true;
        else if (a == "--useCPUReg") o.useCPUReg = true;
        else if (a == "--useGPUReg") o.useGPUReg = true;
        else if (a == "--noStackRegistration") o.noStackRegistration = true;
        else if (a == "--useAutoTemplate") o.useAutoTemplate = true;
        else if (a == "--patchSize") { if (!one(a, v)) return false; o.patchSize = (unsigned)atoi(v.c_str()); }
        else if (a == "--patchStride") { if (!one(a, v)) return false; o.patchStride = (unsigned)atoi(v.c_str()); }
        else if (a == "--disableBiasCorrection") o.disableBiasCorr = true;
        else if (a == "--patchBased") o.patchBased = true;
        else if (a == "--superpixelBased") o.superpixelBased = true;
        else if (a == "--superpixel") { if (!one(a, v)) return false; o.superpixel = (float)atof(v.c_str()); }
        else if (a == "--manualMask") { if (!one(a, o.manualMask)) return false; }
        else if (a == "--useNMI") o.useNMI = true;
        else if (a == "--saveSliceTransformations") o.saveSliceTransformations = true;
        else if (a == "--dump_setup") { if (!one(a, o.dump_setup)) return false; }       // not a reference option: see usage()
        else { std::cerr << "ERROR: unrecognised option '" << argv[i] << "'" << std::endl; return false; }
    }
    if (o.output.empty()) { std::cerr << "ERROR: the option '--output' is required but missing" << std::endl; return false; }
    return true;
}

void write_or_die(const std::string& path, const svr::Image& img)
{
    std::string err;
    if (!svr::write_nifti(path, img, /*as_float32=*/false, &err)) { std::cerr << "cannot write " << path << ": " << err << std::endl; exit(1); }
}

}  // namespace

int main(int argc, char** argv)
{
    Options o;
    bool help = false;
    if (!parse(argc, argv, o, help)) { usage(); return EXIT_FAILURE; }
    if (help) { usage(); return EXIT_SUCCESS; }
    if (o.useCPU) { std::cerr << "FATAL ERROR: this build has no CPU reconstruction path (--useCPU)." << std::endl; return EXIT_FAILURE; }
    if (o.patchBased || o.superpixelBased) { std::cerr << "FATAL ERROR: patch/superpixel-based reconstruction is PVRreconstructionGPU's path." << std::endl; return EXIT_FAILURE; }
    if (!o.packages.empty() || o.T1PackageSize > 0 || !o.sfolder.empty() || !o.manualMask.empty() || o.useAutoTemplate) {
        std::cerr << "FATAL ERROR: --packages / --T1PackageSize / --sfolder / --manualMask / --useAutoTemplate are not supported by this build." << std::endl;
        return EXIT_FAILURE;
    }
    if (!o.useGPUReg) std::cout << "Slice-to-volume and stack registration: the reference's default IRTK rigid registration (cross-correlation, 3 levels), on the GPU." << std::endl;

    std::cout << "Reconstructed volume name ... " << o.output << std::endl;
    size_t nStacks = o.input.size();
    std::cout << "Number of stacks ... " << nStacks << std::endl;
    if (nStacks == 0) { std::cerr << "ERROR: no input stacks (-i)" << std::endl; return EXIT_FAILURE; }

    svr::Image referenceVolume;
    if (!o.referenceVolume.empty()) {
        std::string err; int fr = 1;
        if (!svr::read_nifti(o.referenceVolume, referenceVolume, &fr, &err)) { std::cerr << err << std::endl; return EXIT_FAILURE; }
        std::cout << "using " << o.referenceVolume << " as initial reference volume for " << o.output << std::endl;
    }

    // ---- read the stacks; 4D files are split into one stack per frame (reconstruction.cc:270-320) ------------------
    std::vector<svr::Image> stacks;
    std::vector<double> thickness;
    for (size_t i = 0; i < nStacks; ++i) {
        svr::Image stack; int frames = 1; std::string err;
        if (!svr::read_nifti(o.input[i], stack, &frames, &err)) { std::cerr << "cannot read " << o.input[i] << ": " << err << std::endl; return EXIT_FAILURE; }
        std::cout << "Reading stack ... " << o.input[i] << std::endl;
        if (frames > 1) {
            svr::ImageAttr attr = stack.a;
            attr.z = stack.a.z / frames;
            const size_t n = (size_t)attr.x * attr.y * attr.z;
            for (int t = 0; t < frames; ++t) {
                std::cout << "Splitting stack ... " << o.input[i] << std::endl;
                svr::Image f(attr);
                std::copy(stack.v.begin() + t * n, stack.v.begin() + (t + 1) * n, f.v.begin());
                stacks.push_back(f);
                if (!o.thickness.empty()) thickness.push_back(o.thickness[i]);
            }
        } else {
            stacks.push_back(stack);
            if (!o.thickness.empty()) thickness.push_back(o.thickness[i]);
        }
    }
    char buffer[256];
    for (size_t i = 0; i < stacks.size(); ++i) { snprintf(buffer, sizeof(buffer), "stack%zu.nii", i); write_or_die(buffer, stacks[i]); }
    nStacks = stacks.size();

    // ---- stack transformations (reconstruction.cc:329-353); the template is the first 'id' --------------------------
    int templateNumber = -1;
    std::vector<svr::Rigid> stack_transformations;
    for (size_t i = 0; i < nStacks; ++i) {
        svr::Rigid r;
        if (!o.transformation.empty()) {
            const std::string& t = i < o.transformation.size() ? o.transformation[i] : std::string("id");
            if (t == "id") { if (templateNumber < 0) templateNumber = (int)i; }
            else if (!r.read_dof(t)) { std::cerr << "cannot read transformation " << t << std::endl; return EXIT_FAILURE; }
        } else if (templateNumber < 0) templateNumber = 0;
        stack_transformations.push_back(r);
    }

    // -d d0 d1 ...: one rank per device (one host thread each), slices sharded, NCCL all-reduce of the volume accumulator
    // (reference: reconstruction.cc:355-386 hands the device list to class Reconstruction, cuda2.cu:616-706)
    svr::Reconstruction reconstruction(o.devices.empty() ? std::vector<int>(1, 0) : o.devices);
    if (o.devices.size() > 1) std::cout << "Using " << o.devices.size() << " GPUs (one rank per device)." << std::endl;
    reconstruction.debug = o.debug || o.debug_gpu;
    for (auto& t : stack_transformations) t.invert();          // InvertStackTransformations, irtkReconstructionGPU.cc:5308-5317

    svr::Image maskImage;
    bool have_mask = false;
    if (!o.mask.empty()) {
        std::string err; int fr = 1;
        if (!svr::read_nifti(o.mask, maskImage, &fr, &err)) { std::cerr << "cannot read " << o.mask << ": " << err << std::endl; return EXIT_FAILURE; }
        have_mask = true;
    }
    if (o.num_stacks_tuner > 0) {
        nStacks = o.num_stacks_tuner;
        std::cout << "actually used stacks for tuner test .... " << o.num_stacks_tuner << std::endl;
        stacks.resize(nStacks);
        stack_transformations.resize(nStacks);
    }
    if (thickness.empty()) {
        std::cout << "Slice thickness is ";
        for (size_t i = 0; i < nStacks; ++i) { thickness.push_back(stacks[i].a.dz * 2); std::cout << thickness[i] << " "; }
        std::cout << "." << std::endl;
    }
    reconstruction.SetForceExcludedSlices(o.force_exclude);
    if (templateNumber < 0) { std::cerr << "Please identify the template by assigning id transformation." << std::endl; return EXIT_FAILURE; }
    if (!have_mask) {                               // CreateMask: binarise the template (irtkReconstructionGPU.cc:735-748)
        maskImage = stacks[templateNumber];
        for (double& v : maskImage.v) v = v > 0.0 ? 1 : 0;
        have_mask = true;
        write_or_die("generatedMask.nii.gz", maskImage);
    }

    PerfStats stats;
    stats.start();

    {   // crop the template stack with the mask (reconstruction.cc:566-607)
        svr::Image m = maskImage;
        reconstruction.TransformMask(stacks[templateNumber], m, stack_transformations[templateNumber]);
        reconstruction.CropImage(stacks[templateNumber], m);
        if (o.debug) { write_or_die("maskTemplate.nii.gz", m); write_or_die("croppedTemplate.nii.gz", stacks[templateNumber]); }
    }
    const double resolution = reconstruction.CreateTemplate(stacks[templateNumber], o.resolution);
    reconstruction.SetMask(&maskImage, o.smooth_mask);

    std::streambuf* strm_buffer = std::cout.rdbuf();
    std::streambuf* strm_buffer_e = std::cerr.rdbuf();
    std::ofstream file((o.log_prefix + "log-registration.txt").c_str());
    std::ofstream file_e((o.log_prefix + "log-registration-error.txt").c_str());
    std::ofstream file2((o.log_prefix + "log-reconstruction.txt").c_str());
    std::ofstream fileEv((o.log_prefix + "log-evaluation.txt").c_str());
    std::cout << std::setprecision(3);
    std::cerr << std::setprecision(3);

    // volumetric registration of the stacks to the template (reconstruction.cc:651-662), on the device engine
    const bool stack_registration = o.T1PackageSize == 0 && o.sfolder.empty() && !o.noStackRegistration;
    if (!o.no_log) { std::cerr.rdbuf(file_e.rdbuf()); std::cout.rdbuf(file.rdbuf()); }
    if (stack_registration) reconstruction.StackRegistrations(stacks, stack_transformations, templateNumber);
    if (!o.no_log) { std::cout.rdbuf(strm_buffer); std::cerr.rdbuf(strm_buffer_e); }

    for (size_t i = 0; i < nStacks; ++i) {         // crop the other stacks with the transformed mask (reconstruction.cc:686-707)
        if ((int)i == templateNumber) continue;
        svr::Image m = reconstruction.GetMask();
        reconstruction.TransformMask(stacks[i], m, stack_transformations[i]);
        reconstruction.CropImage(stacks[i], m);
        if (o.debug) { snprintf(buffer, sizeof(buffer), "cropped%zu.nii.gz", i); write_or_die(buffer, stacks[i]); }
    }

    // "Repeat volumetric registrations with cropped stacks" (reconstruction.cc:702-714)
    if (!o.no_log) { std::cerr.rdbuf(file_e.rdbuf()); std::cout.rdbuf(file.rdbuf()); }
    if (stack_registration) reconstruction.StackRegistrations(stacks, stack_transformations, templateNumber);
    if (!o.no_log) { std::cout.rdbuf(strm_buffer); std::cerr.rdbuf(strm_buffer_e); }
    stats.sample("StackRegistrations");

    reconstruction.MatchStackIntensitiesWithMasking(stacks, stack_transformations, o.average, !o.intensity_matching);
    reconstruction.CreateSlicesAndTransformations(stacks, stack_transformations, thickness);
    reconstruction.MaskSlices();
    if (!o.tfolder.empty()) reconstruction.ReadTransformation(o.tfolder);
    (void)resolution;

    stats.sample("overhead/setup");
    if (!o.dump_setup.empty()) {
        if (!o.no_log) { std::cout.rdbuf(strm_buffer); std::cerr.rdbuf(strm_buffer_e); }
        reconstruction.DumpSetup(o.dump_setup);
        std::cout << "setup written to " << o.dump_setup << " (" << reconstruction.NumberOfSlices() << " slices)" << std::endl;
        return EXIT_SUCCESS;
    }
    const auto tick = std::chrono::steady_clock::now();

    reconstruction.SyncGPU();
    if (o.useGPUReg) reconstruction.PrepareRegistrationSlices();
    stats.sample("SyncGPU");
    reconstruction.InitializeEMGPU();
    stats.sample("InitializeEM");
    reconstruction.UpdateGPUTranformationMatrices();

    const int iterations = o.iterations, levels = o.levels;
    for (int iter = 0; iter < iterations; ++iter) {
        if (!o.no_log) std::cout.rdbuf(strm_buffer);
        std::cout << "Iteration " << iter << ". " << std::endl;

        if (iter > 0 || !o.referenceVolume.empty()) {
            if (!o.no_log) { std::cerr.rdbuf(file_e.rdbuf()); std::cout.rdbuf(file.rdbuf()); }
            std::cout << "Iteration " << iter << ": " << std::endl;
            std::cout << "Slice To Volume Registration " << ": " << std::endl;
            if (o.useGPUReg) {                       // reconstruction.cc:868-879
                printf("Slice To Volume Registration GPU\n");
                std::cout << "Slice To Volume Registration GPU" << ": " << std::endl;
                reconstruction.SliceToVolumeRegistrationGPU();
            } else {
                std::cout << "Slice To Volume Registration (IRTK rigid registration, device engine)" << ": " << std::endl;
                reconstruction.SliceToVolumeRegistration();
            }
            stats.sample("Registration");
            std::cout << std::endl;
            if (!o.no_log) std::cerr.rdbuf(strm_buffer_e);
        }

        if (!o.no_log) std::cout.rdbuf(file2.rdbuf());
        std::cout << std::endl << std::endl << "Iteration " << iter << ": " << std::endl << std::endl;

        // smoothing schedule, reconstruction.cc:900-911
        if (iter == iterations - 1) reconstruction.SetSmoothingParameters(o.delta, o.lastIterLambda);
        else {
            double l = o.lambda;
            for (int i = 0; i < levels; ++i) {
                if (iter == iterations * (levels - i - 1) / levels) reconstruction.SetSmoothingParameters(o.delta, l);
                l *= 2;
            }
        }
        reconstruction.generatePSFVolume();
        stats.sample("generatePSFVolume");
        reconstruction.InitializeEMValuesGPU();
        stats.sample("InitializeEMValues");
        reconstruction.UpdateGPUTranformationMatrices();
        stats.sample("CoeffInit");
        reconstruction.GaussianReconstructionGPU();
        {
            reconstruction.SyncCPU();
            snprintf(buffer, sizeof(buffer), "GaussianReconstruction_GPU%i.nii", iter);
            write_or_die(buffer, reconstruction.GetReconstructedGPU());
        }
        stats.sample("GaussianReconstruction");
        reconstruction.SimulateSlicesGPU();
        stats.sample("SimulateSlices");
        reconstruction.InitializeRobustStatisticsGPU();
        stats.sample("InitializeRS");
        reconstruction.EStepGPU();
        stats.sample("EStep");

        const int rec_iterations = iter == iterations - 1 ? (int)o.rec_iterations_last : (int)o.rec_iterations_first;
        for (int i = 0; i < rec_iterations; ++i) {
            std::cout << std::endl << "  Reconstruction iteration " << i << ". " << std::endl;
            if (o.intensity_matching) {
                reconstruction.ScaleGPU();
                stats.sample("Bias and Scale");
            }
            reconstruction.SuperresolutionGPU(i + 1);
            stats.sample("Superresolution");
            if (o.intensity_matching) stats.sample("NormaliseBias");
            reconstruction.SimulateSlicesGPU();
            stats.sample("SimulateSlices");
            reconstruction.MStepGPU(i + 1);
            stats.sample("MStep");
            reconstruction.EStepGPU();
            stats.sample("EStep");
            if (o.debug || o.debug_gpu) {
                reconstruction.SyncCPU();
                snprintf(buffer, sizeof(buffer), "superGPU%i.nii", i);
                write_or_die(buffer, reconstruction.GetReconstructedGPU());
            }
            printf("%d ", i);
        }
        printf("Main loop end\n");
        reconstruction.MaskVolumeGPU();
        stats.sample("MaskVolume");
        printf("Masking done\n");

        reconstruction.SyncCPU();
        stats.sample("SyncCPU");
        snprintf(buffer, sizeof(buffer), "image%i_GPU.nii.gz", iter);
        write_or_die(buffer, reconstruction.GetReconstructedGPU());

        if (o.saveSliceTransformations) reconstruction.SaveTransformations();

        if (!o.no_log) std::cout.rdbuf(fileEv.rdbuf());
        reconstruction.EvaluateGPU(iter, std::cout);
        std::cout << std::endl;
        if (!o.no_log) std::cout.rdbuf(strm_buffer);
        printf("\n");
    }

    reconstruction.RestoreSliceIntensitiesGPU();
    stats.sample("RestoreSliceInt.");
    reconstruction.ScaleVolumeGPU();
    stats.sample("ScaleVolume");
    reconstruction.SyncCPU();
    stats.sample("SyncCPU");

    const double mss = std::chrono::duration<double>(std::chrono::steady_clock::now() - tick).count();
    snprintf(buffer, sizeof(buffer), "performance_GPU_%s.txt", currentDateTime().c_str());
    std::ofstream perf_file(buffer);
    stats.print(std::cout);
    stats.print(perf_file);
    perf_file << "\n.........overall time: " << mss << " s........\n";
    perf_file.close();
    printf(".........overall time: %f s........\n", mss);

    write_or_die(o.output, reconstruction.GetReconstructedGPU());
    return EXIT_SUCCESS;
}
