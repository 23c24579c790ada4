// Internal glue between the C++ host API and the C-ABI render core (include/ptb.h).
#ifndef PTB_HOST_DEVICE_H
#define PTB_HOST_DEVICE_H

#include <PathTrace/base.h>
#include <PathTrace/scene/material.h>
#include <PathTrace/scene/object.h>
#include <ptb.h>

#include <string>

namespace ptb::host {

    //! process-wide context of the calling thread's device ($PTB_DEVICE, else $LOCAL_RANK, else 0), created on first
    //! use; throws std::runtime_error if no CUDA device is usable (there is no CPU fallback)
    ptb_context *defaultContext();

    //! process-wide context of device `device` (the default context when that is its device), created on first use
    ptb_context *contextFor(int device);

    //! throws std::runtime_error carrying ptb_last_error() unless status == PTB_OK
    void check(int status, const char *what);

    //! like check(), for noexcept paths: prints the error once per call site and returns false
    bool ok(int status, const char *what) noexcept;

    //! POD form of a primitive; false for user subclasses of Object
    bool lowerObject(const Object &object, ptb_prim &out) noexcept;

    //! POD form of a (Material, BSDF) pair evaluated at `pos`; false for user BSDF subclasses
    bool lowerMaterial(const Material &material, const BSDF &bsdf, vec3<float> pos, ptb_material &out) noexcept;

}

#endif
