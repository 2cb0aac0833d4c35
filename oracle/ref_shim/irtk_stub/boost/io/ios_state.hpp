#pragma once
#include <ios>
namespace boost { namespace io { struct ios_all_saver { template <class S> explicit ios_all_saver(S&) {} }; } }
