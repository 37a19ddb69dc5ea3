/* Export macro and library initialiser of the B200CVT Graphite plugin (pattern: plugins/OGF/WarpDrive/common/common.h). */
#ifndef H_OGF_B200CVT_COMMON_COMMON_H
#define H_OGF_B200CVT_COMMON_COMMON_H

#include <OGF/basic/common/common.h>

#ifdef B200CVT_EXPORTS
#   define B200CVT_API GEO_EXPORT
#else
#   define B200CVT_API GEO_IMPORT
#endif

namespace OGF {
    static class B200CVT_API B200CVT_libinit {
    public:
        B200CVT_libinit() { increment_users(); }
        ~B200CVT_libinit() { decrement_users(); }
        static void increment_users() { if(count_++ == 0) { initialize(); } }
        static void decrement_users() { if(--count_ == 0) { terminate(); } }
    private:
        static void initialize();
        static void terminate();
        static int count_;
    } B200CVT_libinit_instance;
}
#endif
