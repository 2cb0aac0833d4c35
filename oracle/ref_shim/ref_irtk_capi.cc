// ref_irtk_capi.cc -- TEST / BASELINE INFRASTRUCTURE.  C entry points that forward to the reference's OWN host code: the vendored IRTK
// (source/IRTKSimple2: images, NIfTI I/O, resampling, blurring, rigid transformations, irtkImageRigidRegistrationWithPadding) and
// class irtkReconstruction (source/reconstructionGPU2/irtkReconstructionGPU.cc: set-up pipeline, stack / slice registration, the
// CPU (--useCPU) reconstruction path), all compiled UNMODIFIED where they lie under /root/reference into
// oracle/_ref/libref_irtk.so (oracle/Makefile: make ref_irtk) behind the stand-ins of oracle/ref_shim/irtk/ for the absent GSL, TBB
// and Boost.  The GPU class `Reconstruction` is stubbed out (ref_irtk_stubs.S): this library is the reference's CPU side only and
// runs without a GPU.  No reference code is contained here: every function below is a call into it.
#include <irtkReconstructionGPU.h>
#include <irtkResampling.h>
#include <irtkRegistration.h>
#include <irtkImageRigidRegistration.h>
#include <irtkImageRigidRegistrationWithPadding.h>
#include <irtkImageFunction.h>
#include <irtkTransformation.h>
#include <irtkResamplingWithPadding.h>
#include <irtkGaussianBlurringWithPadding.h>
#include <cstring>
#include <string>
#include <vector>

typedef irtkRealImage Img;

static irtkImageAttributes attrs_from(const double* a)
{
    irtkImageAttributes at;
    at._x = (int)a[0]; at._y = (int)a[1]; at._z = (int)a[2]; at._t = 1;
    at._dx = a[3]; at._dy = a[4]; at._dz = a[5]; at._dt = 1;
    at._xorigin = a[6]; at._yorigin = a[7]; at._zorigin = a[8]; at._torigin = 0;
    for (int i = 0; i < 3; ++i) { at._xaxis[i] = a[9 + i]; at._yaxis[i] = a[12 + i]; at._zaxis[i] = a[15 + i]; }
    return at;
}
static void attrs_to(const irtkImageAttributes& at, double* a)
{
    a[0] = at._x; a[1] = at._y; a[2] = at._z; a[3] = at._dx; a[4] = at._dy; a[5] = at._dz;
    a[6] = at._xorigin; a[7] = at._yorigin; a[8] = at._zorigin;
    for (int i = 0; i < 3; ++i) { a[9 + i] = at._xaxis[i]; a[12 + i] = at._yaxis[i]; a[15 + i] = at._zaxis[i]; }
}
static irtkRigidTransformation rigid_from(const double* d)
{
    irtkRigidTransformation t;
    for (int i = 0; i < 6; ++i) t.Put(i, d[i]);
    return t;
}
static void rigid_to(const irtkRigidTransformation& t, double* d) { for (int i = 0; i < 6; ++i) d[i] = t.Get(i); }

// The GPU member `Reconstruction* reconstructionGPU` is created by a stubbed (do-nothing) constructor: give it the all-zero state
// of default-constructed members, so that the inline helpers the host class calls on it (updateStackSizes: a vector assignment)
// operate on valid empty containers.
struct RecoAccess : public irtkReconstruction {
    RecoAccess(std::vector<int> dev, bool useCPUReg) : irtkReconstruction(dev, useCPUReg)
    {
        memset((void*)reconstructionGPU, 0, sizeof(Reconstruction));
        reconstructionGPU->_useCPUReg = useCPUReg;
    }
};

struct Ctx {
    irtkReconstruction* r;
    std::vector<irtkRealImage> stacks;
    std::vector<irtkRigidTransformation> stack_t;
    std::vector<double> thickness;
};

extern "C" {

void rirtk_set_threads(int n) { tbb_no_threads = n > 0 ? n : tbb::task_scheduler_init::automatic; if (n > 0) tbb::shim_threads() = n; }

// ---- images ----------------------------------------------------------------------------------------------------------------
void* rirtk_image_read(const char* path) { Img* im = new Img(); im->Read(path); return im; }
void* rirtk_image_new(const double* attrs18, const double* data)
{
    Img* im = new Img(attrs_from(attrs18));
    if (data) memcpy(im->GetPointerToVoxels(), data, sizeof(double) * im->GetNumberOfVoxels());
    return im;
}
void rirtk_image_free(void* p) { delete (Img*)p; }
void rirtk_image_attrs(void* p, double* attrs18) { attrs_to(((Img*)p)->GetImageAttributes(), attrs18); }
long rirtk_image_size(void* p) { return ((Img*)p)->GetNumberOfVoxels(); }
void rirtk_image_data(void* p, double* out) { memcpy(out, ((Img*)p)->GetPointerToVoxels(), sizeof(double) * ((Img*)p)->GetNumberOfVoxels()); }
void rirtk_image_write(void* p, const char* path) { ((Img*)p)->Write(path); }
void rirtk_image_matrices(void* p, double* i2w16, double* w2i16)
{
    irtkMatrix a = ((Img*)p)->GetImageToWorldMatrix(), b = ((Img*)p)->GetWorldToImageMatrix();
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) { i2w16[4 * r + c] = a(r, c); w2i16[4 * r + c] = b(r, c); }
}
// irtkResamplingWithPadding<irtkRealPixel> (image++/src/irtkResamplingWithPadding.cc) on a copy
void* rirtk_image_resample_with_padding(void* p, double dx, double dy, double dz, double padding)
{
    Img* out = new Img(*(Img*)p);
    irtkResamplingWithPadding<irtkRealPixel> res(dx, dy, dz, padding);
    res.SetInput(out); res.SetOutput(out); res.Run();
    return out;
}
void* rirtk_image_blur_with_padding(void* p, double sigma, double padding)
{
    Img* out = new Img(*(Img*)p);
    irtkGaussianBlurringWithPadding<irtkRealPixel> blur(sigma, padding);
    blur.SetInput(out); blur.SetOutput(out); blur.Run();
    return out;
}
// rigid transformation: 6 parameters -> matrix (packages/transformation/src/irtkRigidTransformation.cc)
void rirtk_rigid_matrix(const double* dof6, double* m16)
{
    irtkRigidTransformation t = rigid_from(dof6);
    irtkMatrix m = t.GetMatrix();
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m16[4 * r + c] = m(r, c);
}
void rirtk_rigid_from_matrix(const double* m16, double* dof6)
{
    irtkMatrix m(4, 4);
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m(r, c) = m16[4 * r + c];
    irtkRigidTransformation t; t.PutMatrix(m); rigid_to(t, dof6);
}
void rirtk_dof_write(const double* dof6, const char* path) { irtkRigidTransformation t = rigid_from(dof6); t.irtkTransformation::Write((char*)path); }
int rirtk_dof_read(const char* path, double* dof6)
{
    irtkTransformation* t = irtkTransformation::New((char*)path);
    irtkRigidTransformation* r = dynamic_cast<irtkRigidTransformation*>(t);
    if (!r) return 1;
    rigid_to(*r, dof6); delete t; return 0;
}

// ---- the generic registration engine, as the reference's three call sites configure it ---------------------------------------
// kind 0: StackRegistrations (GuessParameterThickSlices, target padding 0; irtkReconstructionGPU.cc:849-1001 minus the mask / offset
//         handling, which the caller does), kind 1: slice / patch to volume (GuessParameterSliceToVolume(false), target padding -1;
//         irtkReconstructionGPU.cc:2030-2036, patchBased2D3DRegistration.cpp:143-147).  Images are cast to irtkGreyImage exactly as there.
double rirtk_rigid_register(void* target_real, void* source_real, int kind, double* dof6)
{
    irtkGreyImage target = *(Img*)target_real;
    irtkGreyImage source = *(Img*)source_real;
    irtkRigidTransformation t = rigid_from(dof6);
    irtkImageRigidRegistrationWithPadding registration;
    registration.SetInput(&target, &source);
    registration.SetOutput(&t);
    if (kind == 0) { registration.GuessParameterThickSlices(); registration.SetTargetPadding(0); }
    else { registration.GuessParameterSliceToVolume(false); registration.SetTargetPadding(-1); }
    registration.Run();
    rigid_to(t, dof6);
    return registration.last_similarity;
}

// The same registration object opened up for the tests: the images irtkImageRegistrationWithPadding::Initialize(level) prepares
// (blurred, resampled, shifted to >= 0 with padding -1; run-length coded padding decoded back to -1) and the similarity
// irtkImageRigidRegistrationWithPadding::Evaluate() returns for a given transformation at that level.
struct RegProbe : public irtkImageRigidRegistrationWithPadding {
    void init_all() { this->irtkImageRegistration::Initialize(); }
    void init_level(int l) { this->irtkImageRegistrationWithPadding::Initialize(l); }
    double eval() { return this->Evaluate(); }
    irtkGreyImage* tgt() { return _target; }
    irtkGreyImage* src() { return _source; }
};
static Img* grey_to_real(irtkGreyImage* g)
{
    Img* out = new Img(g->GetImageAttributes());
    irtkGreyPixel* p = g->GetPointerToVoxels();
    irtkRealPixel* q = out->GetPointerToVoxels();
    for (int i = 0; i < g->GetNumberOfVoxels(); ++i) q[i] = p[i] < 0 ? -1 : p[i];
    return out;
}
double rirtk_reg_probe(void* target_real, void* source_real, int kind, int level, const double* dof6, void** out_target, void** out_source)
{
    irtkGreyImage target = *(Img*)target_real;
    irtkGreyImage source = *(Img*)source_real;
    irtkRigidTransformation t = rigid_from(dof6);
    RegProbe reg;
    reg.SetInput(&target, &source);
    reg.SetOutput(&t);
    if (kind == 0) { reg.GuessParameterThickSlices(); reg.SetTargetPadding(0); }
    else { reg.GuessParameterSliceToVolume(false); reg.SetTargetPadding(-1); }
    reg.init_all();
    reg.init_level(level);
    const double s = reg.eval();
    if (out_target) *out_target = grey_to_real(reg.tgt());
    if (out_source) *out_source = grey_to_real(reg.src());
    return s;
}

// ---- class irtkReconstruction --------------------------------------------------------------------------------------------------
void* rirtk_create(void)
{
    Ctx* c = new Ctx();
    std::vector<int> dev(1, 0);
    c->r = new RecoAccess(dev, true);
    return c;
}
#define R(c) (((Ctx*)(c))->r)
void rirtk_debug(void* c, int on) { if (on) R(c)->DebugOn(); else R(c)->DebugOff(); }
// stacks held by the context (vector<irtkRealImage> as in reconstruction.cc:272-318), with one transformation and thickness each
int rirtk_add_stack(void* c, void* img, const double* dof6, double thickness)
{
    Ctx* x = (Ctx*)c;
    x->stacks.push_back(*(Img*)img); x->stack_t.push_back(rigid_from(dof6)); x->thickness.push_back(thickness);
    return (int)x->stacks.size() - 1;
}
void* rirtk_get_stack(void* c, int i) { return new Img(((Ctx*)c)->stacks[i]); }
void rirtk_get_stack_dof(void* c, int i, double* dof6) { rigid_to(((Ctx*)c)->stack_t[i], dof6); }
void rirtk_invert_stack_transformations(void* c) { R(c)->InvertStackTransformations(((Ctx*)c)->stack_t); }
double rirtk_create_template(void* c, int stack, double resolution) { return R(c)->CreateTemplate(((Ctx*)c)->stacks[stack], resolution); }
void rirtk_set_mask(void* c, void* mask, double sigma, double threshold) { R(c)->SetMask((Img*)mask, sigma, threshold); }
void* rirtk_get_mask(void* c) { return new Img(R(c)->GetMask()); }
void* rirtk_get_reconstructed(void* c) { return new Img(R(c)->GetReconstructed()); }
void rirtk_set_reconstructed(void* c, void* img) { R(c)->SetReconstructed(*(Img*)img); }
// TransformMask + CropImage of a stack with a copy of the mask, as reconstruction.cc:590-607 does per stack
void rirtk_crop_stack_to_mask(void* c, int stack, void* mask)
{
    Ctx* x = (Ctx*)c;
    Img m = *(Img*)mask;
    R(c)->TransformMask(x->stacks[stack], m, x->stack_t[stack]);
    R(c)->CropImage(x->stacks[stack], m);
}
void rirtk_stack_registrations(void* c, int template_number) { Ctx* x = (Ctx*)c; R(c)->StackRegistrations(x->stacks, x->stack_t, template_number); }
void rirtk_match_stack_intensities_with_masking(void* c, double average, int together)
{
    Ctx* x = (Ctx*)c; R(c)->MatchStackIntensitiesWithMasking(x->stacks, x->stack_t, average, together != 0);
}
void rirtk_create_slices_and_transformations(void* c) { Ctx* x = (Ctx*)c; R(c)->CreateSlicesAndTransformations(x->stacks, x->stack_t, x->thickness); }
void rirtk_mask_slices(void* c) { R(c)->MaskSlices(); }
// feed slices directly (synthetic data sets): SetSlicesAndTransformations (irtkReconstructionGPU.cc:1852-1875)
void rirtk_set_slices(void* c, int n, void** imgs, const double* dofs6, const int* stack_ids, const double* thickness)
{
    std::vector<irtkRealImage> sl; std::vector<irtkRigidTransformation> tr; std::vector<int> ids; std::vector<double> th;
    for (int i = 0; i < n; ++i) { sl.push_back(*(Img*)imgs[i]); tr.push_back(rigid_from(dofs6 + 6 * i)); ids.push_back(stack_ids[i]); th.push_back(thickness[i]); }
    R(c)->SetSlicesAndTransformations(sl, tr, ids, th);
}
int rirtk_num_slices(void* c) { std::vector<irtkRealImage> s; R(c)->GetSlices(s); return (int)s.size(); }
void* rirtk_get_slice(void* c, int i) { std::vector<irtkRealImage> s; R(c)->GetSlices(s); return new Img(s[i]); }
void rirtk_get_transformations(void* c, double* dofs6)
{
    std::vector<irtkRigidTransformation> t; R(c)->GetTransformations(t);
    for (size_t i = 0; i < t.size(); ++i) rigid_to(t[i], dofs6 + 6 * i);
}
void rirtk_set_transformations(void* c, int n, const double* dofs6)
{
    std::vector<irtkRigidTransformation> t;
    for (int i = 0; i < n; ++i) t.push_back(rigid_from(dofs6 + 6 * i));
    R(c)->SetTransformations(t);
}
void rirtk_slice_to_volume_registration(void* c) { R(c)->SliceToVolumeRegistration(); }
void rirtk_set_smoothing_parameters(void* c, double delta, double lambda) { R(c)->SetSmoothingParameters(delta, lambda); }
void rirtk_speedup(void* c, int on) { if (on) R(c)->SpeedupOn(); else R(c)->SpeedupOff(); }
// One step of the CPU (--useCPU) loop of reconstruction.cc:800-1276 by name
int rirtk_cpu_step(void* c, const char* name, int iter)
{
    irtkReconstruction* r = R(c);
    const std::string s = name;
    if (s == "InitializeEM") r->InitializeEM();
    else if (s == "InitializeEMValues") r->InitializeEMValues();
    else if (s == "CoeffInit") r->CoeffInit();
    else if (s == "GaussianReconstruction") r->GaussianReconstruction();
    else if (s == "SimulateSlices") r->SimulateSlices();
    else if (s == "InitializeRobustStatistics") r->InitializeRobustStatistics();
    else if (s == "EStep") r->EStep();
    else if (s == "Scale") r->Scale();
    else if (s == "Superresolution") r->Superresolution(iter);
    else if (s == "MStep") r->MStep(iter);
    else if (s == "MaskVolume") r->MaskVolume();
    else if (s == "RestoreSliceIntensities") r->RestoreSliceIntensities();
    else if (s == "ScaleVolume") r->ScaleVolume();
    else return 1;
    return 0;
}

}  // extern "C"
