// Drop-in for the reader of src/stabilization/flowIO.h: ReadFlowFile with the reference's signature, result and
// exception messages (flowIO.cpp:31-78), on top of the C ABI's vsc_flo_read (one fread of the payload instead
// of one per row).  WriteFlowFile is not on the path (the reference never calls it) and is not provided.
#include <stdexcept>
#include <string>
#include <vector>

#include "vsc/vsc.h"

void ReadFlowFile(std::vector<float>& flow, int& width, int& height, std::string filename)
{
    auto fail = [&](int rc) {
        // same text as the reference: "<message> <filename>"
        throw std::runtime_error(std::string(vsc_error_string(rc)) + " " + filename);
    };
    int rc = vsc_flo_read_header(filename.c_str(), &width, &height);
    if (rc)
        fail(rc);
    flow.resize(static_cast<size_t>(height) * width * 2);
    rc = vsc_flo_read(filename.c_str(), flow.data(), flow.size(), &width, &height);
    if (rc)
        fail(rc);
}
