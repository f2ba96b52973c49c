// TEST INFRASTRUCTURE ONLY (part of oracle/_ref, see Makefile.ref).
// The reference's MetalMaterial, compiled from where it lies: materials/metal.cpp is included into this translation
// unit (instead of being compiled on its own) because the measured copper spectra that are the defaults of "eta" and
// "k" (metal.cpp:82-118) are file-local there; this is the only way to ask the reference for their RGB conversion
// (Spectrum::FromSampled, metal.cpp:121-126) without restating the tables.
#include "materials/metal.cpp"

extern "C" void ref_copper_rgb(float *eta_rgb, float *k_rgb) {
    using namespace pbrt;
    Spectrum n = Spectrum::FromSampled(CopperWavelengths, CopperN, CopperSamples);
    Spectrum k = Spectrum::FromSampled(CopperWavelengths, CopperK, CopperSamples);
    n.ToRGB(eta_rgb);
    k.ToRGB(k_rgb);
}
