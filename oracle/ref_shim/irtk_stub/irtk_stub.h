/* TEST INFRASTRUCTURE (oracle/): inert stand-ins for the IRTK host classes the reference's PVR headers mention
 * (include/patchBasedObject.cuh, patchBasedVolume.cuh, reconConfig.cuh ...), so that the UNMODIFIED PVR CUDA sources
 * compile into oracle/_ref/libref_pvr.so.  The parity harness (ref_pvr_capi.cu) never calls the host code that uses
 * them (stack -> patch enumeration, which needs real IRTK): it fills the device structures itself and calls the
 * reference's kernels.  Every method here is a no-op; nothing here is arithmetic of the path; no reference code. */
#pragma once
#include <vector>
#include <string>
#include <iostream>
#include <cmath>
#include <iomanip>
#define TX 0
#define TY 1
#define TZ 2
#define RX 3
#define RY 4
#define RZ 5
struct irtkMatrix {
  double m[4][4];
  irtkMatrix() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = i == j; }
  irtkMatrix(int, int) : irtkMatrix() {}
  void Initialize(int, int) {}
  void Ident() {}
  double& operator()(int i, int j) { return m[i][j]; }
  double operator()(int i, int j) const { return m[i][j]; }
  void Put(int i, int j, double v) { m[i][j] = v; }
  double Get(int i, int j) const { return m[i][j]; }
  void Invert() {}
  void Print() const {}
  irtkMatrix operator*(const irtkMatrix&) const { return *this; }
};
struct irtkVector { };
struct irtkImageAttributes {
  int _x = 0, _y = 0, _z = 0, _t = 1; double _dx = 1, _dy = 1, _dz = 1, _dt = 1;
  double _xorigin = 0, _yorigin = 0, _zorigin = 0, _torigin = 0; double _xaxis[3] = {1,0,0}, _yaxis[3] = {0,1,0}, _zaxis[3] = {0,0,1};
};
template <class T> struct irtkGenericImage {
  irtkGenericImage() {}
  irtkGenericImage(const irtkImageAttributes&) {}
  irtkGenericImage(int, int, int) {}
  template <class U> irtkGenericImage(const irtkGenericImage<U>&) {}
  template <class U> irtkGenericImage& operator=(const irtkGenericImage<U>&) { return *this; }
  irtkGenericImage& operator=(T) { return *this; }
  int GetX() const { return 0; } int GetY() const { return 0; } int GetZ() const { return 0; }
  double GetXSize() const { return 1; } double GetYSize() const { return 1; } double GetZSize() const { return 1; }
  void GetPixelSize(double*, double*, double*) const {}
  void PutPixelSize(double, double, double) {}
  void GetOrigin(double&, double&, double&) const {}
  void PutOrigin(double, double, double) {}
  irtkImageAttributes GetImageAttributes() const { return irtkImageAttributes(); }
  void Initialize(const irtkImageAttributes&) {}
  irtkMatrix GetWorldToImageMatrix() const { return irtkMatrix(); }
  irtkMatrix GetImageToWorldMatrix() const { return irtkMatrix(); }
  void ImageToWorld(double&, double&, double&) const {}
  void WorldToImage(double&, double&, double&) const {}
  T Get(int, int, int, int = 0) const { return T(); }
  void Put(int, int, int, T) {}
  void Put(int, int, int, int, T) {}
  T& operator()(int, int, int, int = 0) { static T t; return t; }
  T* GetPointerToVoxels(int = 0, int = 0, int = 0, int = 0) { return nullptr; }
  int GetNumberOfVoxels() const { return 0; }
  irtkGenericImage GetRegion(int, int, int, int, int, int) const { return *this; }
  void GetMinMax(T*, T*) const {}
  void Write(const char*) const {}
  void Read(const char*) {}
};
typedef irtkGenericImage<double> irtkRealImage;
typedef irtkGenericImage<short> irtkGreyImage;
typedef double irtkRealPixel; typedef short irtkGreyPixel;
struct irtkTransformation { virtual ~irtkTransformation() {} };
struct irtkRigidTransformation : irtkTransformation {
  irtkMatrix GetMatrix() const { return irtkMatrix(); }
  void PutMatrix(const irtkMatrix&) {}
  void UpdateParameter() {}
  void Invert() {}
  void PutTranslationX(double) {} void PutTranslationY(double) {} void PutTranslationZ(double) {}
  void PutRotationX(double) {} void PutRotationY(double) {} void PutRotationZ(double) {}
  double GetTranslationX() const { return 0; } double GetTranslationY() const { return 0; } double GetTranslationZ() const { return 0; }
  double GetRotationX() const { return 0; } double GetRotationY() const { return 0; } double GetRotationZ() const { return 0; }
  void Transform(double&, double&, double&) const {}
  void Print() const {}
};
